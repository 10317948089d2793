"""Refit a FitSNAP dump on the GPU:  python tools/fit_npy.py Descriptors.npy Truth-Ref.npy Weights.npy
       [--alpha 1e-8] [--chunk-rows 1000000] [--refine 2] [--out coeffs.npy]
The three files are what `[EXTRAS] dump_descriptors / dump_truth / dump_weights = 1` writes
(fitsnap3lib/calculators/calculator.py:329-337).  They are memory-mapped and streamed chunk by chunk
(`fitsnap_b200.pipeline.StreamingLinearFit`), so matrices larger than device or host memory are fine."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("descriptors")
    ap.add_argument("truth")
    ap.add_argument("weights")
    ap.add_argument("--alpha", type=float, default=0.0, help="ridge parameter ([RIDGE] alpha); 0 = least squares (SVD solver)")
    ap.add_argument("--chunk-rows", type=int, default=1 << 20)
    ap.add_argument("--refine", type=int, default=2)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from fitsnap_b200.pipeline import StreamingLinearFit, npy_row_chunks
    chunks = npy_row_chunks(args.descriptors, args.truth, args.weights, args.chunk_rows)
    res = StreamingLinearFit(alpha=args.alpha, refine=args.refine).fit(chunks)
    x = res.coefficients()
    info = res.info_host()
    print("rows %d, coefficients %d, kernel launches %d, factor status %d (dropped columns %d, zero columns %d)"
          % (res.extra["rows_streamed"], x.shape[0], res.launches, info[0], info[3], info[2]))
    if args.out:
        np.save(args.out, x)
    else:
        np.set_printoptions(precision=17, linewidth=120)
        print(x)


if __name__ == "__main__":
    main()
