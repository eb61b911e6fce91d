"""Config objects for the SA-M4C hot path.

`BertConfig.from_dict` mirrors the behaviour `train.py` relies on
(/root/reference/train.py:92-93, pytorch-transformers `BertConfig.from_dict`): start from the
BERT-base defaults, then every key of the YAML section becomes an attribute, used or not.

`c3_config()` restates the model sections of the shipped experiment file
/root/reference/configs/train-tvqa-eval-tvqa-c3.yml:47-88 as plain data, so benches and tests
do not need the reference tree.
"""
import copy


class BertConfig(object):
    _DEFAULTS = dict(
        vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
        intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
        attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
        initializer_range=0.02, layer_norm_eps=1e-12, output_attentions=False,
        output_hidden_states=False, torchscript=False, pruned_heads={},
    )

    def __init__(self, **kwargs):
        for k, v in self._DEFAULTS.items():
            setattr(self, k, copy.deepcopy(v))
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_dict(cls, json_object):
        config = cls()
        for key, value in dict(json_object).items():
            config.__dict__[key] = value
        return config

    def to_dict(self):
        return copy.deepcopy(self.__dict__)

    def get(self, key, default=None):
        return self.__dict__.get(key, default)

    def __repr__(self):
        return "BertConfig(%r)" % (self.__dict__,)


_C3_MMT = dict(
    num_hidden_layers=2, num_spatial_layers=4, heads_type="mix",
    layer_type_list=["n", "n", "s", "s", "s", "s"],
    mix_list=["none", "none", "share3", "share3", "share3", "share3"],
    obj_drop=0.1, ocr_drop=0.1, hidden_size=768, num_spatial_relations=12, type_vocab_size=2,
    vocab_size=30522, textvqa_vocab_size=3998, pooling_method="mul", ptr_query_size=768,
    ocr_feature_size=3002, obj_feature_size=2048, finetune_ocr_obj=False, use_phoc_fasttext=True,
    normalize=True, lr_scale_mmt=1.0, num_decoding_steps=12, max_obj_num=100, max_ocr_num=50,
    max_seq_length=20, beam_size=1, attention_mask_quadrants=[1, 2],
)
_C3_TEXTBERT = dict(lr_scale_text_bert=0.1, num_hidden_layers=3,
                    text_bert_init_from_bert_base=False, vocab_size=30522)


def c3_config(**mmt_overrides):
    """(mmt_dict, text_bert_dict) of the shipped c3 experiment; overrides apply to the MMT section.

    text_bert_init_from_bert_base is False here: the pretrained BERT-base checkpoint is a
    network download in the reference (sa_m4c.py:74-77) and is not available offline.
    """
    mmt = copy.deepcopy(_C3_MMT)
    mmt.update(mmt_overrides)
    return mmt, copy.deepcopy(_C3_TEXTBERT)
