"""TEST INFRASTRUCTURE ONLY (oracle shim) -- never imported by the product path.

The reference pins the third-party package `pytorch-transformers==1.0.0`
(`/root/reference/requirements.txt:1`) and imports nine classes from its `modeling_bert`
module (`/root/reference/sam/sa_m4c.py:8-10, 617, 663-664, 718`).  That package is not
vendored under /root/reference, is not installed in this image, and there is no network.
This file restates the *published semantics* of those nine classes (post-LN BERT block as in
Devlin et al. 2018 / the pytorch-transformers 1.x source) in plain torch so that the
UNMODIFIED reference `sam/sa_m4c.py` can be imported here and used to pin our oracle:

  BertConfig.from_dict        BERT-base defaults, then every dict key becomes an attribute
  BertLayerNorm               TF style: (x-u)/sqrt(var+eps), biased variance, eps 1e-12
  BertEmbeddings              dropout(LN(word + position + token_type)), padding_idx=0
  BertSelfAttention           softmax(QK^T/sqrt(d_h) + additive_mask) -> dropout -> PV
  BertSelfOutput / BertOutput LN(dropout(dense(h)) + residual)
  BertIntermediate            erf-GELU(dense(h))
  BertLayer / BertEncoder     tuple-returning, head_mask argument
  BertPreTrainedModel         init_weights(): N(0, initializer_range) for Linear/Embedding
                              weights, zero Linear bias, LN weight 1 / bias 0

tests/test_oracle_shim.py cross-checks BertLayer/BertEmbeddings against the installed
`transformers` package (same math, different API) so a typo here cannot go unnoticed.
Parameter names match pytorch-transformers exactly (they define the state_dict contract).
"""
import math

import torch
from torch import nn


class BertConfig(object):
    def __init__(self, vocab_size_or_config_json_file=30522, hidden_size=768,
                 num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 hidden_act="gelu", hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512,
                 type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12, **kwargs):
        self.vocab_size = vocab_size_or_config_json_file
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_act = hidden_act
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range
        self.layer_norm_eps = layer_norm_eps
        # PretrainedConfig base attributes
        self.output_attentions = kwargs.pop("output_attentions", False)
        self.output_hidden_states = kwargs.pop("output_hidden_states", False)
        self.torchscript = kwargs.pop("torchscript", False)
        self.pruned_heads = kwargs.pop("pruned_heads", {})

    @classmethod
    def from_dict(cls, json_object):
        config = cls(vocab_size_or_config_json_file=-1)
        for key, value in json_object.items():
            config.__dict__[key] = value
        return config

    def to_dict(self):
        return dict(self.__dict__)


def _gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


ACT2FN = {"gelu": _gelu, "relu": torch.nn.functional.relu}


class BertLayerNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        u = x.mean(-1, keepdim=True)
        s = (x - u).pow(2).mean(-1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.variance_epsilon)
        return self.weight * x + self.bias


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids=None, position_ids=None):
        seq_length = input_ids.size(1)
        if position_ids is None:
            position_ids = torch.arange(seq_length, dtype=torch.long, device=input_ids.device)
            position_ids = position_ids.unsqueeze(0).expand_as(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        e = (self.word_embeddings(input_ids) + self.position_embeddings(position_ids)
             + self.token_type_embeddings(token_type_ids))
        return self.dropout(self.LayerNorm(e))


class BertSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("hidden size not a multiple of the number of heads")
        self.output_attentions = config.output_attentions
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = int(config.hidden_size / config.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def _split(self, x):
        return x.view(x.size(0), x.size(1), self.num_attention_heads,
                      self.attention_head_size).permute(0, 2, 1, 3)

    def forward(self, hidden_states, attention_mask, head_mask=None):
        q = self._split(self.query(hidden_states))
        k = self._split(self.key(hidden_states))
        v = self._split(self.value(hidden_states))
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.attention_head_size)
        scores = scores + attention_mask
        probs = self.dropout(nn.Softmax(dim=-1)(scores))
        if head_mask is not None:
            probs = probs * head_mask
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
        ctx = ctx.view(ctx.size(0), ctx.size(1), self.all_head_size)
        return (ctx, probs) if self.output_attentions else (ctx,)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)

    def forward(self, input_tensor, attention_mask, head_mask=None):
        self_outputs = self.self(input_tensor, attention_mask, head_mask)
        return (self.output(self_outputs[0], input_tensor),) + self_outputs[1:]


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        act = config.hidden_act
        self.intermediate_act_fn = ACT2FN[act] if isinstance(act, str) else act

    def forward(self, hidden_states):
        return self.intermediate_act_fn(self.dense(hidden_states))


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def forward(self, hidden_states, attention_mask, head_mask=None):
        attention_outputs = self.attention(hidden_states, attention_mask, head_mask)
        attention_output = attention_outputs[0]
        layer_output = self.output(self.intermediate(attention_output), attention_output)
        return (layer_output,) + attention_outputs[1:]


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.output_attentions = config.output_attentions
        self.output_hidden_states = config.output_hidden_states
        self.layer = nn.ModuleList([BertLayer(config) for _ in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask, head_mask=None):
        all_hidden, all_attn = (), ()
        for i, layer_module in enumerate(self.layer):
            if self.output_hidden_states:
                all_hidden = all_hidden + (hidden_states,)
            outs = layer_module(hidden_states, attention_mask,
                                None if head_mask is None else head_mask[i])
            hidden_states = outs[0]
            if self.output_attentions:
                all_attn = all_attn + (outs[1],)
        if self.output_hidden_states:
            all_hidden = all_hidden + (hidden_states,)
        outputs = (hidden_states,)
        if self.output_hidden_states:
            outputs = outputs + (all_hidden,)
        if self.output_attentions:
            outputs = outputs + (all_attn,)
        return outputs


class BertPreTrainedModel(nn.Module):
    config_class = BertConfig
    base_model_prefix = "bert"

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        self.config = config

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, BertLayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def init_weights(self):
        self.apply(self._init_weights)

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("oracle shim: no network, set text_bert_init_from_bert_base=false")
