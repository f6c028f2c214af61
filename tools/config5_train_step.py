#!/usr/bin/env python
"""BASELINE configs[4]: one CoarseNet-style training step with this repo's geometry path in the loss, data-parallel over the
GPUs of one box with NCCL gradient all-reduce.  A CALLER of the hot path (a load generator), not part of it:

    render the current prediction (vertices_transform_raw -> rendering_layer: maskimg | pncc | normal, 7 channels, as
    nets/network.py:108-121 feeds CoarseNet) -> stock torchvision ResNet-101 (7-channel stem, 235 outputs, random init,
    bf16 autocast; the reference's regressor is a TF-slim resnet_v1, :122-134) -> raw 235-d prediction ->
    vertices_transform_raw (set_constraints fused, :204-218) -> rendering_layer -> loss = MSE(depth map, target) +
    MSE(constrained params, label) + geometry loss (Gram form, :346-355) -> backward through the rasterizer's and the
    reconstruction's backward kernels into the regressor -> DDP all-reduce (bf16 compression hook) -> Adam step.

    python tools/config5_train_step.py [--batch 64 --steps 10 --warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/config5_train_step.py

Prints one JSON line on rank 0: step time (CUDA events, max over ranks), samples/s over all GPUs, the bytes every rank
all-reduces per step, and the step time with gradient synchronisation switched off (DistributedDataParallel.no_sync) --
the difference is the all-reduce time that the overlap with the backward pass does not hide.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "3dfacerecon_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64, help="faces per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.nn as nn
    import torchvision
    from torch.nn.parallel import DistributedDataParallel as DDP
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    dist_mod = importlib.import_module(PKG + ".distributed")
    net_mod = importlib.import_module(PKG + ".nets.network")
    rank, local_rank, world = dist_mod.init_from_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, S = args.batch, 200
    torch.manual_seed(1234)                                             # same initial weights on every rank

    model = synth.make_synthetic_model(seed=0, jitter=0.2)
    gray = torch.rand((B, S, S, 1), device=dev)
    geo = net_mod.FaceRecNet(im_gray=gray, mesh_data=model, batch_size=B, im_size=S, device=dev)
    regressor = torchvision.models.resnet101(num_classes=geo.ndim)
    regressor.conv1 = nn.Conv2d(7, 64, kernel_size=7, stride=2, padding=3, bias=False)     # maskimg 1 + pncc 3 + normal 3 channels
    nn.init.normal_(regressor.fc.weight, std=1e-3)                      # network.py:131
    nn.init.zeros_(regressor.fc.bias)
    regressor = regressor.to(dev).to(memory_format=torch.channels_last)
    nparams = sum(p.numel() for p in regressor.parameters())
    ddp = None
    if world > 1:
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        ddp = DDP(regressor, device_ids=[local_rank], gradient_as_bucket_view=True)
        ddp.register_comm_hook(None, default_hooks.bf16_compress_hook)
    fwd = ddp if ddp is not None else regressor
    opt = torch.optim.Adam(regressor.parameters(), lr=1e-4)

    rng = np.random.default_rng(100 + rank)
    label = torch.from_numpy(synth.sample_params_constrained(B, seed=200 + rank)).to(dev)
    target_depth = torch.rand((B, S, S, 1), device=dev)
    start_raw = torch.zeros((B, geo.ndim), device=dev)                  # sigmoid(0) = the centre of every constraint range
    del rng

    def step(sync=True):
        opt.zero_grad(set_to_none=True)
        with torch.no_grad():                                           # rendering of the current prediction: the CNN's input
            vp0 = geo.vertices_transform_raw(start_raw)
            pncc, normal, maskimg, _ = geo.rendering_layer(vp0, geo.tri, geo.vertex_code)
            x = torch.cat([maskimg, pncc, normal], dim=3).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        ctx = ddp.no_sync() if (ddp is not None and not sync) else torch.autocast("cuda", enabled=False)
        with ctx:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                raw = fwd(x)
            raw = raw.float()
            vp = geo.vertices_transform_raw(raw)                        # set_constraints fused into the prep kernels
            _, _, _, depthimg = geo.rendering_layer(vp, geo.tri, geo.vertex_code)
            pred = geo.set_constraints(raw[:, None, None, :]).squeeze(2).squeeze(1)
            loss = (nn.functional.mse_loss(depthimg, target_depth) + nn.functional.mse_loss(pred[:, :7], label[:, :7]) +
                    1e-9 * geo.geometry_loss(pred, label))
            loss.backward()
        opt.step()
        return loss

    def timed(sync):
        for _ in range(args.warmup):
            step(sync)
        dist_mod.barrier()
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            loss = step(sync)
        b.record()
        torch.cuda.synchronize(dev)
        dist_mod.barrier()
        return dist_mod.reduce_scalar(a.elapsed_time(b) / args.steps, "max"), float(loss.detach())

    ms_sync, loss = timed(True)
    ms_nosync = timed(False)[0] if ddp is not None else None
    if rank == 0:
        print(json.dumps({
            "workload": "BASELINE configs[4]: CoarseNet-style training step, ResNet-101 regressor (random init, bf16 autocast, 7-channel "
                        "rendered input, 235 outputs) + recon/render loss through this repo's forward and backward kernels",
            "n_gpus": world, "batch_per_gpu": B, "ms_per_step": ms_sync, "samples_per_s": B * world / (ms_sync * 1e-3),
            "regressor_params": nparams, "allreduce_bytes_per_rank_per_step": 2 * nparams if ddp is not None else 0,
            "allreduce": "torch DDP over NCCL, bf16 compression hook, bucketed and overlapped with the backward pass" if ddp is not None else None,
            "ms_per_step_without_grad_sync": ms_nosync,
            "exposed_allreduce_ms": None if ms_nosync is None else ms_sync - ms_nosync,
            "loss": loss, "steps": args.steps, "warmup": args.warmup}))
    dist_mod.shutdown()


if __name__ == "__main__":
    main()
