"""The prompt-encoder oracle (oracle/text_oracle.py) pinned against the REAL upstream modules the RepText pipelines call -
transformers' ``T5EncoderModel`` and ``CLIPTextModel`` (RepText/pipeline_flux_controlnet.py:289-291, :330) - with their
own random initialisation, in this container (CPU, fp32)."""
import pytest
import torch

from oracle import text_oracle as TO
from util import rel_l2

tf = pytest.importorskip("transformers")


def _perturb_norms(m):
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "layer_norm" in n or "layernorm" in n.lower():
                p.add_(0.1 * torch.randn_like(p))


@pytest.mark.parametrize("S", [16, 200])
def test_t5_oracle_matches_transformers(S):
    torch.manual_seed(0)
    cfg = dict(TO.T5_XXL_CONFIG, vocab_size=100, d_model=128, d_ff=256, num_layers=3, num_heads=2)
    hf = tf.T5Config(feed_forward_proj="gated-gelu", dropout_rate=0.0, **cfg)
    m = tf.T5EncoderModel(hf).eval()
    _perturb_norms(m)
    with torch.no_grad():       # the default init of the relative bias is tiny: make it matter
        m.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight.normal_(0, 1.0)
    ids = torch.randint(0, 100, (2, S))
    with torch.no_grad():
        want = m(ids)[0]
        got = TO.t5_encoder(m.state_dict(), cfg, ids)
    assert rel_l2(got, want) < 1e-5
    assert set(TO.t5_param_shapes(cfg)) <= set(m.state_dict())
    for k, s in TO.t5_param_shapes(cfg).items():
        assert tuple(m.state_dict()[k].shape) == s, k


def test_t5_buckets_match_transformers():
    from transformers.models.t5.modeling_t5 import T5Attention
    S = 512
    i = torch.arange(S)
    want = T5Attention._relative_position_bucket(i[None, :] - i[:, None], bidirectional=True, num_buckets=32, max_distance=128)
    lut = TO.t5_relative_buckets(S)
    assert torch.equal(lut[(i[None, :] - i[:, None]) + S - 1], want)


@pytest.mark.parametrize("eos", [2, 7])
def test_clip_oracle_matches_transformers(eos):
    torch.manual_seed(1)
    cfg = dict(TO.CLIP_L_CONFIG, vocab_size=100, hidden_size=128, intermediate_size=256, num_hidden_layers=3,
               num_attention_heads=2, max_position_embeddings=24, eos_token_id=eos)
    hf = tf.CLIPTextConfig(hidden_act="quick_gelu", attention_dropout=0.0, bos_token_id=0, pad_token_id=1, **cfg)
    m = tf.CLIPTextModel(hf).eval()
    _perturb_norms(m)
    ids = torch.randint(8, 99, (3, 24))
    ids[:, 0] = 0
    for b, pos in enumerate((5, 17, 23)):          # EOS (the largest id / eos id) somewhere in each row, padding after
        ids[b, pos] = 99 if eos == 2 else eos
        ids[b, pos + 1:] = 1
    with torch.no_grad():
        out = m(ids)
        last, pooled = TO.clip_text(m.state_dict(), cfg, ids)
    assert rel_l2(last, out.last_hidden_state) < 1e-5
    assert rel_l2(pooled, out.pooler_output) < 1e-5
    for k, s in TO.clip_param_shapes(cfg).items():
        assert tuple(m.state_dict()[k].shape) == s, k
