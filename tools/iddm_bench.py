"""IDDM class-conditional sampler throughput: DDIMDiffusion.sample(model, n, labels, cfg_scale) at
64x64 with classifier-free guidance, 50 and 500 sampling steps (500 is the reference's default, utils/initializer.py:169),
16-bit and fp32 modes.  One JSON line per point; CUDA events around one call after a warm-up call."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--steps", default="50,500")
    ap.add_argument("--precisions", default="bf16,fp32")
    args = ap.parse_args()
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import iddm
    torch.manual_seed(0)
    net = iddm.UNet(num_classes=37, image_size=args.size).eval().cuda()
    labels = torch.randint(0, 37, (args.n,), device="cuda")
    for precision in args.precisions.split(","):
        net.set_precision(precision)
        for steps in map(int, args.steps.split(",")):
            ddim = iddm.DDIMDiffusion(noise_steps=1000, sample_steps=steps, img_size=args.size, device="cpu")
            ddim.sample(net, args.n, labels=labels, cfg_scale=3)          # warm-up: engine build + graph capture
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            img = ddim.sample(net, args.n, labels=labels, cfg_scale=3)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            eng = net.engine(2 * args.n)
            print(json.dumps({"path": "iddm", "precision": "16-bit" if precision == "bf16" else precision, "images": args.n,
                              "size": args.size, "sample_steps": len(ddim.time_step), "cfg_scale": 3,
                              "images_per_s": round(args.n / (ms / 1e3), 2), "ms": round(ms, 1),
                              "attention": sorted(set(eng.attn_kinds)), "launches_per_step": len(eng.L) + 4,
                              "out": [str(img.dtype), list(img.shape)]}), flush=True)
        net.release_engines()


if __name__ == "__main__":
    main()
