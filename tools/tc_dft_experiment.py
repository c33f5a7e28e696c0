"""Tensor-core experiment: a DFT stage of the 38192-point search transform evaluated as a dense
contraction on the tensor cores, measured instead of argued.

The contraction is run through cuBLAS (torch.matmul) -- a library GEMM is the *best case* for a hand-written tcgen05
kernel of the same shape (it already uses tcgen05.mma + TMA with tuned tiles), and it is used here only as an
experiment, never on the product path.  Shapes:

  * the whole 217-point pass (R = 31 x 7) as one complex DFT matrix:   [434 x 434] x [434 x 176 B]
  * only the 31-point stage (the O(p^2) butterfly of the CUDA-core path): [62 x 62] x [62 x 1232 B]

Complex arithmetic is embedded in real GEMMs ([[Re, -Im], [Im, Re]]).  Precisions: fp16 with a two-term split of both
operands (hi*hi + hi*lo + lo*hi, three GEMMs, ~22 mantissa bits -- what the 1e-5 peak-metric parity needs), single
fp16/bf16 (one GEMM, ~1e-3), tf32.  Output: JSON lines with time per transform and the error against float64.
"""
import json
import sys
import numpy as np
import torch


def dft_real(p, inverse=True):
    k = np.arange(p)
    w = np.exp((2j if inverse else -2j) * np.pi * np.outer(k, k) / p)
    return np.block([[w.real, -w.imag], [w.imag, w.real]])


def split16(x, dt):
    hi = x.to(dt)
    lo = (x - hi.to(torch.float32)).to(dt)
    return hi, lo


def mm32(a, b):
    """fp16/bf16 operands, fp32 accumulator AND fp32 output (what a TMEM accumulator read back with tcgen05.ld gives)."""
    return torch.mm(a, b, out_dtype=torch.float32)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1676          # transforms per launch of the CUDA-core pass (ncu_summary_r1_v5)
    dev = "cuda"
    torch.backends.cuda.matmul.allow_fp16_reduced_precision_reduction = False      # fp32 accumulation, as a TMEM accumulator
    torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = False
    rng = np.random.default_rng(1)
    for name, p, cols_per_tr in (("pass217", 217, 176), ("stage31", 31, 1232)):
        # zero-padded to multiples of 64 / 128: with the raw 434 (or 62) cuBLAS falls back to an sm_80-style mma.sync kernel
        # (cutlass_80_tensorop_s16816gemm, align2); padded, it picks its sm_100 (tcgen05) kernels
        q = (2 * p + 63) // 64 * 64
        W64 = np.zeros((q, q))
        W64[:2 * p, :2 * p] = dft_real(p)
        n = (cols_per_tr * B + 127) // 128 * 128
        X64 = np.zeros((q, min(n, 65536)))
        X64[:2 * p] = rng.standard_normal((2 * p, X64.shape[1]))      # error check on a slice
        Y64 = W64 @ X64
        W = torch.tensor(W64, dtype=torch.float32, device=dev)
        X = torch.randn((q, n), dtype=torch.float32, device=dev)
        X[2 * p:] = 0
        X[:, :X64.shape[1]] = torch.tensor(X64, dtype=torch.float32, device=dev)
        flop = 2.0 * (2 * p) * (2 * p) * cols_per_tr * B              # useful work (the padding is not counted)
        res = []
        # fp32 on the CUDA cores (cuBLAS sgemm), for scale
        torch.backends.cuda.matmul.allow_tf32 = False
        ms, Y = timed(lambda: W @ X)
        res.append(("fp32 sgemm", 1, ms, Y))
        torch.backends.cuda.matmul.allow_tf32 = True
        ms, Y = timed(lambda: W @ X)
        res.append(("tf32", 1, ms, Y))
        torch.backends.cuda.matmul.allow_tf32 = False
        for dt, nm in ((torch.float16, "fp16"), (torch.bfloat16, "bf16")):
            sc = 1.0 / 8.0                                       # keep |x| well inside the fp16 range
            Wh, Wl = split16(W, dt)
            Xh, Xl = split16(X * sc, dt)
            ms, Y = timed(lambda: mm32(Wh, Xh) / sc)
            res.append((nm + " x1", 1, ms, Y))

            def three():
                acc = mm32(Wh, Xh)
                acc += mm32(Wh, Xl)
                acc += mm32(Wl, Xh)
                return acc / sc
            ms, Y = timed(three)
            res.append((nm + " split x3 (incl. fp32 adds of the three products)", 3, ms, Y))
            # the three products alone (what a fused kernel accumulating in TMEM would pay)
            ms, _ = timed(lambda: (mm32(Wh, Xh), mm32(Wh, Xl), mm32(Wl, Xh)))
            res.append((nm + " split x3 (GEMMs only)", 3, ms, Y))
        for label, terms, ms, Y in res:
            err = float(np.abs(Y[:, :X64.shape[1]].double().cpu().numpy() - Y64).max() / np.abs(Y64).max())
            print(json.dumps({"shape": name, "precision": label, "transforms": B, "ms": round(ms, 4),
                              "us_per_transform": round(ms * 1e3 / B, 4), "tflops": round(terms * flop / ms / 1e9, 1),
                              "max_err_rel_to_max": err}))


if __name__ == "__main__":
    main()
