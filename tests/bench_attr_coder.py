"""(Lives under tests/ because it loads the reference extension from oracle/_ref: only tests/, smoke() and bench.py may touch oracle/.)
SURVEY 8f-4: throughput of the attribute coder on one B200 -- the library's fused Gaussian path and its table path against the
reference extension (oracle/_ref/arithmetic.so, built from the reference's own sources) on the same tensors.

    python tests/bench_attr_coder.py [n_symbols] [reps]       (default 10 000 000 = one file of encoder_gaussian_chunk, 5 reps)

Prints one JSON line.  Times are CUDA-event times around the calls the reference's encoder_gaussian / decoder_gaussian make
(calculate_cdf + arithmetic_encode, calculate_cdf + arithmetic_decode); bytes / symbols are checked for identity first."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gauspcc_b200 import arithmetic as A

CHUNK = 10000


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    mean = torch.randn(n, device=dev, generator=g) * 3
    scale = torch.randn(n, device=dev, generator=g).abs() * 0.5 + 0.3
    Q = torch.full((n,), 0.25, device=dev)
    x = mean + scale * torch.randn(n, device=dev, generator=g)
    xi = torch.round(x / Q)
    mn, mx = int(xi.min()), int(xi.max())
    Lp = mx - mn + 2
    sym = (xi - mn).to(torch.int16)

    t_enc, (stream, cnt) = timed(lambda: A.gaussian_encode(sym, mean, scale, Q, mn, mx, CHUNK), reps)
    t_dec, dec = timed(lambda: A.gaussian_decode(mean, scale, Q, stream, cnt, mn, mx, CHUNK), reps)
    assert torch.equal(dec, sym)

    def enc_table():
        return A.arithmetic_encode(sym, A.calculate_cdf(mean, scale, Q, mn, mx), CHUNK, n, Lp)

    def dec_table():
        return A.arithmetic_decode(A.calculate_cdf(mean, scale, Q, mn, mx), stream, cnt, CHUNK, n, Lp)

    t_enc_t, (s2, c2) = timed(enc_table, reps)
    t_dec_t, d2 = timed(dec_table, reps)
    assert torch.equal(s2, stream) and torch.equal(c2, cnt) and torch.equal(d2, sym)
    line = {"metric": "Msymbols/s HAC attribute coder (calculate_cdf + arithmetic_encode / _decode, device-timed)",
            "n_symbols": n, "Lp": Lp, "chunk_size": CHUNK, "bits_per_symbol": round(8 * stream.numel() / n, 3),
            "ours_fused": {"enc_ms": round(t_enc, 3), "dec_ms": round(t_dec, 3),
                           "enc_Msym_s": round(n / t_enc / 1e3, 1), "dec_Msym_s": round(n / t_dec / 1e3, 1)},
            "ours_table": {"enc_ms": round(t_enc_t, 3), "dec_ms": round(t_dec_t, 3)}}
    from oracle import build_ref_arithmetic
    ref = build_ref_arithmetic.load_module()
    if ref is None:
        line["reference"] = "oracle/_ref/arithmetic.so not built"
    else:
        def enc_ref():
            return ref.arithmetic_encode(sym, ref.calculate_cdf(mean, scale, Q, mn, mx), CHUNK, n, Lp)

        def dec_ref():
            return ref.arithmetic_decode(ref.calculate_cdf(mean, scale, Q, mn, mx), stream, cnt, CHUNK, n, Lp)

        t_enc_r, (s3, c3) = timed(enc_ref, max(1, reps // 2))
        t_dec_r, d3 = timed(dec_ref, max(1, reps // 2))
        line["reference"] = {"enc_ms": round(t_enc_r, 3), "dec_ms": round(t_dec_r, 3), "identical_stream": bool(torch.equal(s3, stream) and torch.equal(c3, cnt)),
                             "identical_symbols": bool(torch.equal(d3, sym))}
        line["speedup_vs_reference_ext"] = {"enc": round(t_enc_r / t_enc, 1), "dec": round(t_dec_r / t_dec, 1)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
