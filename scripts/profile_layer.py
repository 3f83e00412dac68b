"""Run single binarized layers of the ResNet-18 workload in isolation (for ncu / event timing).

    python scripts/profile_layer.py [--layers l1,l2,l3,l4,...] [--reps 5] [--batch 256] [--flags 0]
Prints one JSON object per layer with CUDA-event times of the pack and conv launches."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bnn_b200  # noqa: E402
from bnn_b200 import functional as BF  # noqa: E402

LAYERS = {  # name: (c_in, c_out, h, w, k, stride, pad)
    "l1": (64, 64, 56, 56, 3, 1, 1), "l2s": (64, 128, 56, 56, 3, 2, 1), "l2": (128, 128, 28, 28, 3, 1, 1),
    "l2d": (64, 128, 28, 28, 1, 1, 0), "l3s": (128, 256, 28, 28, 3, 2, 1), "l3": (256, 256, 14, 14, 3, 1, 1),
    "l3d": (128, 256, 14, 14, 1, 1, 0), "l4s": (256, 512, 14, 14, 3, 2, 1), "l4": (512, 512, 7, 7, 3, 1, 1),
    "l4d": (256, 512, 7, 7, 1, 1, 0),
    # ResNet-50 (use --batch 128): the 1x1 expand convs (conv3: residual + fp32 out + planes) and reduce convs (conv1: planes only)
    "r50_l1c3": (64, 256, 56, 56, 1, 1, 0), "r50_l2c3": (128, 512, 28, 28, 1, 1, 0), "r50_l3c3": (256, 1024, 14, 14, 1, 1, 0),
    "r50_l4c3": (512, 2048, 7, 7, 1, 1, 0), "r50_l2c1": (512, 128, 28, 28, 1, 1, 0), "r50_l3c1": (1024, 256, 14, 14, 1, 1, 0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default="l1,l2,l3,l4")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--fused", default="", help="'mid' (bn+relu -> bits only) or 'out' (bn+res+relu -> fp32 + bits)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    for name in args.layers.split(","):
        ci, co, h, w, k, s, p = LAYERS[name]
        x = torch.relu(torch.randn(args.batch, ci, h, w, device=dev))
        wt = torch.randn(co, ci, k, k, device=dev) * 0.05
        wts = BF.pack_weights(wt, True, True)
        act = BF.pack_activations(x)
        out = BF.bconv2d(act, wts, None, None, (s, s), (p, p), (1, 1), flags=args.flags)
        bn = (torch.rand(co, device=dev) + 0.5, torch.randn(co, device=dev) * 0.2)
        res = torch.randn_like(out)
        res_cl = res.contiguous(memory_format=torch.channels_last)

        def run_conv():
            if args.fused == "mid":
                BF.bconv2d_fused(act, wts, bn=bn, activation=1, want_out=False, want_bits=True, stride=(s, s),
                                 padding=(p, p), flags=args.flags)
            elif args.fused == "out_cl":
                BF.bconv2d_fused(act, wts, bn=bn, residual=res_cl, activation=1, want_out=True, want_bits=True,
                                 stride=(s, s), padding=(p, p), flags=args.flags, channels_last=True)
            elif args.fused == "out":
                BF.bconv2d_fused(act, wts, bn=bn, residual=res, activation=1, want_out=True, want_bits=True,
                                 stride=(s, s), padding=(p, p), flags=args.flags)
            else:
                BF.bconv2d(act, wts, None, None, (s, s), (p, p), (1, 1), flags=args.flags, out=out)

        run_conv()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tp = tc = 0.0
        for _ in range(args.reps):
            ev[0].record()
            act = BF.pack_activations(x)
            ev[1].record()
            run_conv()
            ev[2].record()
            torch.cuda.synchronize()
            tp += ev[0].elapsed_time(ev[1]) / args.reps
            tc += ev[1].elapsed_time(ev[2]) / args.reps
        ho, wo = out.shape[2], out.shape[3]
        bmac = args.batch * co * ho * wo * ci * k * k
        print(json.dumps({"layer": name, "fused": args.fused, "pack_ms": round(tp, 4), "conv_ms": round(tc, 4),
                          "tbmac_s": round(bmac / tc * 1e-9, 1),
                          "out_gb_s": round(out.numel() * 4 / tc * 1e-6, 1)}), flush=True)


if __name__ == "__main__":
    main()
