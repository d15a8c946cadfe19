"""Host logic of the prompt-encoder drop-ins (reptext_b200/text_encoders.py) WITHOUT a GPU: the operator entry points are
replaced by the torch statement of their contracts (tests/ops_emulator.py), fp32, and compared with the oracle - which
tests/test_text_oracle.py pins against transformers.  Covers the segment layout of the fused q|k|v and wi_0|wi_1 launches,
the relative-bias look-up table indexed by (key - query), the EOS pooling and the error behaviour."""
import pytest
import torch

import ops_emulator
from oracle import text_oracle as TO
from reptext_b200 import ops, text_encoders as TE
from reptext_b200.models import FrozenConfig
from util import rel_l2


def _cpu(cls, defaults, cfg, sd):
    m = object.__new__(cls)                   # the constructor insists on a CUDA device; the emulator runs on the CPU
    full = dict(defaults)
    full.update(cfg)
    m.config = FrozenConfig(**full)
    m.dtype, m.device, m._w = torch.float32, torch.device("cpu"), dict(sd)
    m._bias_lut = {}
    return m


def test_t5_host_logic(monkeypatch):
    ops_emulator.install(monkeypatch, ops)
    cfg = dict(TO.T5_XXL_CONFIG, vocab_size=100, d_model=128, d_ff=256, num_layers=2, num_heads=2)
    sd = TO.random_state_dict(TO.t5_param_shapes(cfg), seed=1)
    m = _cpu(TE.T5EncoderModel, TE.T5_XXL_CONFIG, cfg, sd)
    ids = torch.randint(0, 100, (2, 40), generator=torch.Generator().manual_seed(2))
    want = TO.t5_encoder(sd, cfg, ids)
    out = m(ids, output_hidden_states=False)
    assert rel_l2(out[0], want) < 1e-5 and out.last_hidden_state is out[0]
    assert torch.equal(TE.t5_relative_buckets(512, 32, 128), TO.t5_relative_buckets(512))
    with pytest.raises(IndexError):
        m(torch.full((1, 8), 100))
    with pytest.raises(ValueError):
        m(ids, attention_mask=torch.ones(2, 40))
    assert TE.t5_param_shapes(cfg) == TO.t5_param_shapes(cfg)


@pytest.mark.parametrize("eos", [2, 7])
def test_clip_host_logic(monkeypatch, eos):
    ops_emulator.install(monkeypatch, ops)
    cfg = dict(TO.CLIP_L_CONFIG, vocab_size=100, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
               num_attention_heads=2, max_position_embeddings=24, eos_token_id=eos)
    sd = TO.random_state_dict(TO.clip_param_shapes(cfg), seed=3)
    m = _cpu(TE.CLIPTextModel, TE.CLIP_L_CONFIG, cfg, sd)
    ids = torch.randint(8, 99, (3, 24), generator=torch.Generator().manual_seed(4))
    for b, pos in enumerate((5, 17, 23)):
        ids[b, pos] = 99 if eos == 2 else eos
        ids[b, pos + 1:] = 1
    last, pooled = TO.clip_text(sd, cfg, ids)
    out = m(ids)
    assert rel_l2(out.last_hidden_state, last) < 1e-5
    assert rel_l2(out.pooler_output, pooled) < 1e-5
    with pytest.raises(ValueError):
        m(torch.zeros(1, 25, dtype=torch.long))
    assert TE.clip_param_shapes(cfg) == TO.clip_param_shapes(cfg)


def test_text_encoders_reject_cpu_and_odd_configs():
    with pytest.raises(ValueError):
        TE.T5EncoderModel({}, {}, device="cpu")
    with pytest.raises(ValueError):
        TE.T5EncoderModel(dict(d_kv=128), {}, device="cuda")
    with pytest.raises(ValueError):
        TE.CLIPTextModel(dict(hidden_size=512, num_attention_heads=4), {}, device="cuda")
    with pytest.raises(ValueError):
        TE.CLIPTextModel({}, {}, dtype=torch.float32, device="cuda")
