"""One T5-XXL (512 tokens) and one CLIP-L (77 tokens) prompt encode with random weights after a warm-up: the program ncu
profiles for the prompt-encoder launch list; prints the per-class device time (option "profile")."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reptext_b200 import _lib, text_encoders as TE  # noqa: E402
from reptext_b200.pipeline_utils import SyntheticTokenizer  # noqa: E402

dev = torch.device("cuda")
t5 = TE.T5EncoderModel(None, TE.random_weights(TE.t5_param_shapes(TE.T5_XXL_CONFIG), 103, dev))
clip = TE.CLIPTextModel(None, TE.random_weights(TE.clip_param_shapes(TE.CLIP_L_CONFIG), 104, dev))
p = "a street sign in city, with the text 'مرحبا بالعالم', filmfotos, film grain"
i5 = SyntheticTokenizer("t5", 32128, 512)([p], padding="max_length", max_length=512, truncation=True).input_ids.to(dev)
ic = SyntheticTokenizer("clip", 49408, 77)([p], padding="max_length", max_length=77, truncation=True).input_ids.to(dev)
t5(i5), clip(ic)
torch.cuda.synchronize()
for name, fn in (("t5_xxl", lambda: t5(i5)), ("clip_l", lambda: clip(ic))):
    for simt in (0, 1, 2):     # 0: tcgen05 attention (the product), 1: CUDA cores, 2: warp-level mma.sync
        _lib.set_option("text_attn_simt", int(simt == 1))
        _lib.set_option("text_attn_mma", int(simt == 2))
        _lib.set_option("profile", 1)
        _lib.profile_reset()
        fn()
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.set_option("profile", 0)
        _lib.profile_reset()
        rec = {k: dict(ms=round(v[0], 3), launches=v[2],
                       achieved=round(v[1] / (v[0] / 1e3) / (1e9 if k in ("elementwise",) else 1e12), 2),
                       unit="GB/s" if k in ("elementwise",) else "TFLOP/s") for k, v in prof.items()}
        print(json.dumps({name + ("", "_simt_attention", "_mma_sync_attention")[simt]: rec}), flush=True)
_lib.set_option("text_attn_simt", 0)
_lib.set_option("text_attn_mma", 0)
