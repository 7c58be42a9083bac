"""BASELINE config 5: the simulator feeding the reference's physics-informed model, end to end.

    python examples/train_e2e.py [--reference baseline/_ref] [--samples 64] [--batched 64] [--iters 20]

Uses the REFERENCE's own dataset class (src/utils/data_loader.py: SyntheticSmokeDataset) and model
(src/models/smokephys_net.py: SmokePhysNet) from an installed copy of the reference, with `src.physics`
resolved to this repo's CUDA simulator (smokephysai_b200.run.merge_src).  The generation phase runs twice:
through the reference's unmodified per-sample loop (scalar class surface) and through the batched back-end
(SmokeSimulator.generate_dataset); then a few AdamW iterations of the model on the generated frames.
The DataLoader is built here with num_workers=0: the reference's create_data_loaders forks workers over a
dataset of CUDA tensors, which fails in stock PyTorch independently of the simulator.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.path.join(ROOT, "baseline", "_ref"))
    ap.add_argument("--samples", type=int, default=64)
    ap.add_argument("--batched", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--input-dim", type=int, default=32, help="SmokePhysNet token grid (config.yaml uses 128: 16k tokens)")
    a = ap.parse_args()

    import numpy as np
    import torch
    import torch.nn.functional as F
    from torch.utils.data import DataLoader
    from smokephysai_b200.run import merge_src, patch_batched_generation
    merge_src(a.reference)
    import src.utils.data_loader as dl
    from src.models.smokephys_net import SmokePhysNet

    dev = torch.device("cuda")
    out = {"samples": a.samples, "grid": [128, 128], "sequence_length": 20}

    np.random.seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ds_loop = dl.SyntheticSmokeDataset(num_samples=a.samples, grid_size=(128, 128), device=dev)        # reference loop, our simulator
    torch.cuda.synchronize()
    out["generation_reference_loop_s"] = time.perf_counter() - t0

    patch_batched_generation(a.batched)
    np.random.seed(0)
    t0 = time.perf_counter()
    ds = dl.SyntheticSmokeDataset(num_samples=a.samples, grid_size=(128, 128), device=dev)             # batched back-end
    torch.cuda.synchronize()
    out["generation_batched_s"] = time.perf_counter() - t0
    out["generation_batched_cell_steps_per_s"] = a.samples * 20 * 128 * 128 / out["generation_batched_s"]
    same = all(torch.equal(x["sequence"], y["sequence"]) for x, y in zip(ds_loop.data, ds.data))
    out["batched_equals_loop"] = bool(same)

    loader = DataLoader(ds, batch_size=8, shuffle=True, num_workers=0)
    model = SmokePhysNet(input_dim=a.input_dim, hidden_dim=128, num_layers=2, num_heads=4).to(dev)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    model.train()
    it, losses = 0, []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    while it < a.iters:
        for batch in loader:
            opt.zero_grad()
            o = model(batch["input"].to(dev))
            loss = F.mse_loss(o["reconstructed"], batch["target"].to(dev)) + 0.1 * F.mse_loss(o["physics_features"], batch["chaos_features"].to(dev))
            loss.backward()
            opt.step()
            losses.append(loss.item())
            it += 1
            if it >= a.iters:
                break
    torch.cuda.synchronize()
    out["train_iters"] = it
    out["train_s"] = time.perf_counter() - t0
    out["loss_first"], out["loss_last"] = losses[0], losses[-1]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
