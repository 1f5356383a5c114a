import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rag_gesture_b200 as R
from rag_gesture_b200 import _lib, config as C, synthetic as S, ops
dev = torch.device("cuda:0")
model = R.build_submodule(dict(C.denoiser_cfg(), precision=_lib.PREC_FP32), database=None, use_retrieval_for_test=False)
model.load_state_dict(S.synthetic_state_dict(0), strict=False)
model = model.to(dev).eval()
blk = model.temporal_decoder_blocks[0].sa_block
B, T = 3, 43
torch.manual_seed(0)
for scale in (1.0, 5.0):
    h = torch.randn(B, T, 512, device=dev) * scale
    w = torch.cat([blk.query.weight, blk.key.weight, blk.value.weight], 0)
    b = torch.cat([blk.query.bias, blk.key.bias, blk.value.bias], 0)
    qkv = ops.linear(ops.layernorm(h, blk.norm.weight, blk.norm.bias), w, b)
    for mname, mask in (("ones", torch.ones(B, T, device=dev)), ("sep", S.motion_mask(B).to(dev))):
        y0 = ops.self_attention_core(qkv, mask, 0)
        for mode in (1, 2):
            y = ops.self_attention_core(qkv, mask, mode)
            d = (y - y0).abs()
            print(f"scale {scale} mask {mname} mode {mode}: max|d| {d.max().item():.3e} ref max {y0.abs().max().item():.3e} "
                  f"rel-L2 {(d.norm() / y0.norm()).item():.3e}  worst idx {divmod(d.view(B*T, 512).argmax().item(), 512)}")
    print("qkv stats: q", qkv[..., :512].abs().max().item(), "k", qkv[..., 512:1024].abs().max().item(), "v", qkv[..., 1024:].abs().max().item())
