"""Control-node warp: csrc/warp.cu vs the same statements in torch on the same GPU.

The torch arm is the restatement oracle/g4r_oracle.py control_node_warp_ref moved to the device, with torch.topk on the [N, M]
distance matrix standing in for pytorch3d.ops.knn_points (absent here) -- i.e. what the reference's ControlNodeWarp.forward
launches after the node MLP, up to the knn kernel.  fwd + bwd of sum(outputs * fixed gradients), CUDA events."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "4dgs-slam_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def torch_arm(x, nodes, lr, wl, attrs, mask, K):
    n3 = nodes[..., :3].detach()
    dist = torch.cdist(x, n3) ** 2
    nn_dist, nn_idx = torch.topk(dist, K, dim=1, largest=False)
    w = torch.exp(-nn_dist / (2 * torch.exp(lr)[nn_idx] ** 2)) * torch.sigmoid(wl)[nn_idx][..., 0] + 1e-7
    w = w / w.sum(-1, keepdim=True)
    bias = torch.tensor([1.0, 0, 0, 0], device=x.device)
    q = attrs["local_rotation"] + bias
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    R = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r), two_s * (i * j + k * r), 1 - two_s * (i * i + k * k),
                     two_s * (j * k - i * r), two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1).reshape(-1, 3, 3)
    nn_nodes = n3[nn_idx]
    Ax = torch.einsum("nkab,nkb->nka", R[nn_idx], x[:, None] - nn_nodes) + nn_nodes + attrs["d_xyz"][nn_idx]
    translate = ((Ax * w[..., None]).sum(1) - x) * mask
    rotation = (attrs["d_rotation"][nn_idx] * w[..., None]).sum(1) * mask
    scale = (attrs["d_scaling"][nn_idx] * w[..., None]).sum(1) * mask
    return translate, rotation, scale


def main():
    from diff_gaussian_rasterization.deform import control_node_warp
    dev = torch.device("cuda", 0)
    rows = {}
    for N, M, K in ((30_000, 512, 3), (150_000, 512, 3), (500_000, 512, 3)):
        g = torch.Generator(device="cpu").manual_seed(0)
        rn = lambda *s: torch.randn(*s, generator=g).to(dev)
        x, nodes = rn(N, 3), rn(M, 3)
        lr, wl = (rn(M) * 0.3 - 0.5).requires_grad_(), rn(M, 1).requires_grad_()
        attrs = {k: (rn(M, c) * 0.3).requires_grad_() for k, c in (("d_xyz", 3), ("d_rotation", 4), ("d_scaling", 3), ("local_rotation", 4))}
        mask = torch.ones(N, 1, device=dev)
        gT, gR, gS = rn(N, 3), rn(N, 4), rn(N, 3)
        params = [lr, wl] + list(attrs.values())

        def ours():
            o = control_node_warp(x, nodes, lr, wl, attrs, mask, K=K)
            torch.autograd.grad((o["d_xyz"] * gT).sum() + (o["d_rotation"] * gR).sum() + (o["d_scaling"] * gS).sum(), params)

        def theirs():
            t, r, s = torch_arm(x, nodes, lr, wl, attrs, mask, K)
            torch.autograd.grad((t * gT).sum() + (r * gR).sum() + (s * gS).sum(), params)

        def time_ms(fn, iters=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        rows[f"N{N}_M{M}_K{K}"] = dict(ms_ours=time_ms(ours), ms_torch=time_ms(theirs))
        rows[f"N{N}_M{M}_K{K}"]["speedup"] = rows[f"N{N}_M{M}_K{K}"]["ms_torch"] / rows[f"N{N}_M{M}_K{K}"]["ms_ours"]
        print(N, rows[f"N{N}_M{M}_K{K}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "deform_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
