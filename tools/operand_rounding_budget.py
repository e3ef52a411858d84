"""Which 16-bit tensors of the path cost how much map accuracy (CPU emulation, no GPU needed).

    python tools/operand_rounding_budget.py [img_size] [seed]

The engine keeps the residual stream, LayerNorm statistics, softmax statistics and all accumulation in fp32; what it
rounds to 16 bits are the tensors that feed the tensor cores: the raw residual copy x16, q|k|v, the softmax
probabilities P, the attention output, the GELU output, the ln_post tokens, the patch operand, and every GEMM weight
(in_proj and c_fc after the LayerNorm gamma has been folded in). This script restates the path in fp32 PyTorch with
those rounding points as switches and reports the max abs error of the level-summed anomaly map against the all-fp32
run (tolerance of the task: 1e-2) for: everything in bf16 / fp16, ONE class in bf16 (its own contribution), and
everything in bf16 EXCEPT one class (what keeping that class wider would buy).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import aaclip_b200  # noqa: F401
from aaclip_b200 import weights

S = int(sys.argv[1]) if len(sys.argv) > 1 else 224
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
CLASSES = ["patchA", "x16", "qkv", "P", "att", "hid", "tok", "W_patch", "W_qkv", "W_out", "W_fc", "W_proj", "W_adapter",
           "W_seg"]


def run(sd, images, text, mode):
    """mode: dict class -> torch dtype (rounded) ; missing = fp32."""
    def r(t, c):
        d = mode.get(c)
        return t if d is None else t.to(d).float()

    p = "image_encoder."
    W = 1024
    x = F.unfold(images, 14, stride=14).transpose(1, 2)                       # [B,P,588]
    x = r(x, "patchA") @ r(sd[p + "conv1.weight"].reshape(W, -1), "W_patch").t()
    cls = sd[p + "class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], 1) + sd[p + "positional_embedding"]
    x = F.layer_norm(x, (W,), sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"], 1e-5)
    taps = []

    def folded(x, g, b, Wt, bias, wc):
        """LN(x) Wt^T + bias the way the engine computes it: raw 16-bit x, W' = Wt * gamma rounded once."""
        mean, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
        rstd = (var + 1e-5).rsqrt()
        Wp = r(Wt * g, wc)
        acc = r(x, "x16") @ Wp.t()
        return rstd * (acc - mean * Wp.sum(1)) + (bias + Wt @ b)

    for i in range(24):
        b = f"{p}transformer.resblocks.{i}."
        qkv = r(folded(x, sd[b + "ln_1.weight"], sd[b + "ln_1.bias"], sd[b + "attn.in_proj_weight"],
                       sd[b + "attn.in_proj_bias"], "W_qkv"), "qkv")
        B, N, _ = qkv.shape
        q, k, v = (t.view(B, N, 16, 64).transpose(1, 2) for t in qkv.chunk(3, -1))
        s = (q @ k.transpose(-1, -2)) * 0.125
        m = s.max(-1, keepdim=True).values
        pexp = torch.exp(s - m)
        att = (r(pexp, "P") @ v) / pexp.sum(-1, keepdim=True)                 # fp32 row sum, 16-bit P (as the kernel)
        att = r(att.transpose(1, 2).reshape(B, N, W), "att")
        x = x + att @ r(sd[b + "attn.out_proj.weight"], "W_out").t() + sd[b + "attn.out_proj.bias"]
        h = F.gelu(folded(x, sd[b + "ln_2.weight"], sd[b + "ln_2.bias"], sd[b + "mlp.c_fc.weight"],
                          sd[b + "mlp.c_fc.bias"], "W_fc"))
        x = x + r(h, "hid") @ r(sd[b + "mlp.c_proj.weight"], "W_proj").t() + sd[b + "mlp.c_proj.bias"]
        if i < 6:
            a = F.leaky_relu(r(x, "x16") @ r(sd[f"image_adapter.layer_adapters.{i}.fc.0.weight"], "W_adapter").t(), 0.01)
            a = a * x.norm(dim=-1, keepdim=True) / a.norm(dim=-1, keepdim=True)
            x = 0.1 * a + 0.9 * x
        if i + 1 in (6, 12, 18, 24):
            taps.append(F.layer_norm(x[:, 1:], (W,), sd[p + "ln_post.weight"], sd[p + "ln_post.bias"], 1e-5))
    total = 0
    G = S // 14
    for l, t in enumerate(taps):
        f = F.normalize(r(t, "tok") @ r(sd[f"image_adapter.seg_proj.{l}.fc.weight"], "W_seg").t(), dim=-1)
        sc = 100.0 * f @ text
        total = total + ((sc[..., 1] + 1 - sc[..., 0]) / 2).view(-1, 1, G, G)
    return F.interpolate(total, size=S, mode="bilinear", align_corners=True)   # (the blur is linear: left out)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    sd = weights.make_visual_state_dict(seed=seed, img_size=S)
    images, text = weights.make_images(seed, 1, S), weights.make_text_anchor(seed)
    with torch.no_grad():
        ref = run(sd, images, text, {})
        err = lambda mode: (run(sd, images, text, mode) - ref).abs().max().item()
        print(f"# {S} px, seed {seed}; max abs error of the level-summed map (before the blur) vs all-fp32; tolerance 1e-2")
        print(f"all 16-bit tensors fp16 : {err({c: torch.float16 for c in CLASSES}):.2e}")
        all_bf = {c: torch.bfloat16 for c in CLASSES}
        e_all = err(all_bf)
        print(f"all 16-bit tensors bf16 : {e_all:.2e}")
        print(f"{'class':10s} {'only this in bf16':>18s} {'all bf16 but this (fp32)':>26s} {'all bf16 but this (fp16)':>26s}")
        for c in CLASSES:
            only = err({c: torch.bfloat16})
            but32 = err({k: v for k, v in all_bf.items() if k != c})
            but16 = err({**all_bf, c: torch.float16})
            print(f"{c:10s} {only:18.2e} {but32:26.2e} {but16:26.2e}")
        acts = ["patchA", "x16", "qkv", "P", "att", "hid", "tok"]
        print(f"activations fp16, weights bf16 : {err({c: (torch.float16 if c in acts else torch.bfloat16) for c in CLASSES}):.2e}")
        print(f"activations bf16, weights fp16 : {err({c: (torch.bfloat16 if c in acts else torch.float16) for c in CLASSES}):.2e}")


if __name__ == "__main__":
    main()
