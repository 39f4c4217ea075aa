"""Per-stage timing of one Jacobi forward inside the persistent kernel (run on a GPU box):
    python tools/stage_profile.py [--model qwen2.5-coder-7b] [--forward 3] [--gen 128]
Prints the mean microseconds per stage kind (work of CTA 0 + wait at the grid barrier) and writes the raw rows."""
import argparse
import json
import os
import random
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="qwen2.5-coder-7b")
    ap.add_argument("--forward", type=int, default=3)
    ap.add_argument("--gen", type=int, default=128)
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--prompt", type=int, default=512)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--mode", default="single")
    ap.add_argument("--tp-attn", default="auto", choices=["auto", "shard", "replicate"])
    ap.add_argument("--attn-chunk", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "stage_profile.json"))
    a = ap.parse_args()
    import torch
    from jacobiforcing_b200.runtime import JacobiRuntime, hf_generate
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(rank)
    group = None
    if world > 1:                                   # tensor-parallel profile: torchrun --nproc-per-node tp
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
        group = dist.group.WORLD
    cfg = bench.SHAPES[a.model]
    pages = (a.prompt + a.gen + 3 * a.n + 255) // 256 + 1
    rt = JacobiRuntime(cfg, device=rank, max_rows=max(a.prompt + a.n, 64, a.batch * a.n * 5), max_seqs=a.batch,
                       num_kv_pages=pages * a.batch, max_pages_per_seq=pages, attn_chunk=a.attn_chunk, gen_capacity=a.prompt + a.gen + 4 * a.n + 8,
                       tp_rank=rank, tp_size=world, tp_group=group, tp_shard_attention={'auto': None, 'shard': True, 'replicate': False}[a.tp_attn])
    bench.load_random_weights(rt, cfg, 1234, f"cuda:{rank}")
    g = torch.Generator().manual_seed(1)
    prompts = [torch.randint(0, cfg["vocab_size"], (a.prompt,), generator=g).tolist() for _ in range(a.batch)]
    for rep in range(2):
        rt.profile_enable(a.forward if rep == 1 else -1)
        seqs, results, run = hf_generate(rt, prompts, n=a.n, max_new_tokens=a.gen, mode=a.mode, pad=0, K=2, r=0.85,
                                         n_gram_pool_size=4, rngs=[random.Random(2 + i) for i in range(a.batch)])
        for s in seqs:
            rt.release(s)
    if rank != 0:
        return
    prof = rt.profile_read()
    kinds = defaultdict(lambda: [0.0, 0.0, 0])
    total = 0.0
    for row in prof["stages"]:
        k = row["stage"].split(".")[-1]
        kinds[k][0] += row["work_us"]; kinds[k][1] += row["wait_us"]; kinds[k][2] += 1
        total += row["work_us"] + row["wait_us"]
    print(f"SM clock during the profiled forward: {prof.get('sm_clock_khz_measured', 0) / 1e3:.0f} MHz (attribute: {prof['sm_clock_khz'] / 1e3:.0f})")
    print(f"forward {a.forward}: {total:.1f} us over {len(prof['stages'])} stages; run: {run['forwards']} forwards, "
          f"{run['device_ms']:.2f} ms -> {run['device_ms'] * 1e3 / max(1, run['forwards']):.1f} us / forward")
    print(f"{'stage':<20}{'count':>6}{'work_us(mean)':>16}{'wait_us(mean)':>16}{'total_us':>12}{'share':>8}")
    for k, (wk, wt, cnt) in sorted(kinds.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
        print(f"{k:<20}{cnt:>6}{wk / cnt:>16.2f}{wt / cnt:>16.2f}{wk + wt:>12.1f}{(wk + wt) / total:>8.1%}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump({"model": a.model, "forward": a.forward, "run": {k: v for k, v in run.items() if k != "trace"},
               "profile": prof}, open(a.out, "w"))


if __name__ == "__main__":
    main()
