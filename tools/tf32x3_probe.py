"""Is a 3xTF32 split (hi*hi + lo*hi + hi*lo on the tensor cores, FP32 accumulate) precise enough for K1?

Builds the sliding windows of a synthetic feature file explicitly, computes cross(t) = sum_c sum_i q[c][i] b[c][t+i]
(a) in FP64, (b) with FP32 torch matmul (no TF32), (c) with three TF32 tensor-core matmuls on split operands, for the
whole K = 14*172 chain in ONE accumulator and for per-channel accumulators summed in FP32 afterwards.  Errors are
reported relative to the correlation norm, i.e. as errors of the final sim."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from strugatzki_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
mu, sigma, floor0, norm = synth.default_profile(14)
T, W = 20000, 172
raw = synth.synth_file(synth.BASE_SEED, 3, T + W, mu, sigma, floor0)
b = ((raw - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32)          # [frames][14]
q = synth.synth_file(synth.BASE_SEED, 0, W, mu, sigma, floor0)
q = ((q - norm[:, 0]) / (norm[:, 1] - norm[:, 0])).astype(np.float32)
# a planted near-copy so that some sims are ~1
b[5000:5000 + W] = q + 0.01 * np.random.default_rng(1).standard_normal(q.shape).astype(np.float32)
qs = q[:, 1:]
taps = (qs.astype(np.float64) - qs.astype(np.float64).mean()).astype(np.float32)   # spectral group, zero mean
bs = torch.from_numpy(b[:, 1:]).to(dev)                                           # [frames][13]
win = bs.unfold(0, W, 1)[:T]                                                      # [T][13][W] view
A = win.reshape(T, 13 * W).contiguous()                                           # explicit windows (c, i)
x = torch.from_numpy(taps.T.copy()).to(dev).reshape(13 * W)                       # (c, i)
A64, x64 = A.double(), x.double()
ref = A64 @ x64
# sim normalisation: cross / (sqrt(sum q~^2) * sqrt(sum (b - mean_b)^2))
nq = torch.sqrt((x64 * x64).sum())
nb = torch.sqrt(((A64 - A64.mean(1, keepdim=True)) ** 2).sum(1))
den = nq * nb
sim = ref / den


def split(v):
    hi = (v.view(torch.int32) & ~0x1FFF).view(torch.float32)
    lo = v - hi
    lo_hi = (lo.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, lo_hi


out = {"T": T, "K": 13 * W, "max_sim": float(sim.max())}
X = x.reshape(-1, 1).repeat(1, 8).contiguous()
torch.backends.cuda.matmul.allow_tf32 = False
fp32 = (A @ X)[:, 0]
out["fp32_matmul_max_sim_err"] = float(((fp32.double() - ref) / den).abs().max())
torch.backends.cuda.matmul.allow_tf32 = True
torch.set_float32_matmul_precision("high")
Ah, Al = split(A)
Xh, Xl = split(X)
one = (A @ X)[:, 0]
out["tf32x1_max_sim_err"] = float(((one.double() - ref) / den).abs().max())
three = (Ah @ Xh + (Al @ Xh + Ah @ Xl))[:, 0]
out["tf32x3_one_chain_max_sim_err"] = float(((three.double() - ref) / den).abs().max())
out["tf32x3_one_chain_max_rel_err_big"] = float((((three.double() - ref) / den).abs() / sim.abs())[sim.abs() > 0.3].max())
# per-channel accumulators (K = 172 each), summed in FP32
acc = torch.zeros(T, device=dev)
for c in range(13):
    sl = slice(c * W, (c + 1) * W)
    acc += (Ah[:, sl] @ Xh[sl] + (Al[:, sl] @ Xh[sl] + Ah[:, sl] @ Xl[sl]))[:, 0]
out["tf32x3_per_channel_max_sim_err"] = float(((acc.double() - ref) / den).abs().max())
e = ((acc.double() - ref) / den)
out["tf32x3_per_channel_mean_sim_err"] = float(e.mean())
out["tf32x3_per_channel_rms_sim_err"] = float(e.pow(2).mean().sqrt())
e1 = ((three.double() - ref) / den)
out["tf32x3_one_chain_mean_sim_err"] = float(e1.mean())
out["tf32x3_one_chain_rms_sim_err"] = float(e1.pow(2).mean().sqrt())
e0 = ((fp32.double() - ref) / den)
out["fp32_rms_sim_err"] = float(e0.pow(2).mean().sqrt())
print(json.dumps(out))
