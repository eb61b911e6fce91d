"""TEST INFRASTRUCTURE ONLY -- import stand-in for pytorch_transformers.tokenization_bert (requirements.txt:1).
sam/task_utils.py:7 imports BertTokenizer at module level; it is used by load_datasets only (off the hot path)."""


class BertTokenizer(object):
    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        raise RuntimeError("no tokenizer files in this environment (oracle shim)")
