"""Host-side tensor-parallel logic on CPU: the shard function reassembles to the full tensors and the sharded
forward arithmetic (column-parallel -> row-parallel with a sum over ranks) equals the unsharded one; exercised with a
real 2-process gloo group as the transport, the way the runtime exchanges IPC handles."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jacobiforcing_b200.tp import check_divisibility, shard_tensor


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.hf_model import build_model
        model = build_model("tiny", seed=0, dtype=torch.float32)
        sd = model.state_dict()
        ok = True
        # 1. shards gathered over the group reassemble every tensor
        for name, t in sd.items():
            sh = shard_tensor(name, t, rank, world).contiguous()
            parts = [torch.empty_like(sh) for _ in range(world)]
            dist.all_gather(parts, sh)
            from jacobiforcing_b200.tp import shard_dim
            dim, _ = shard_dim(name)
            full = torch.cat(parts, dim=dim) if (dim >= 0 and not (t.dim() == 1 and dim == 1) and sh.shape != t.shape) else parts[0]
            ok = ok and torch.equal(full, t)
        # 2. one MLP + attention-projection block computed tensor-parallel == unsharded (sum over ranks = all_reduce)
        layer = model.model.layers[0]
        x = torch.randn(5, model.config.hidden_size, generator=torch.Generator().manual_seed(3))
        p = "model.layers.0."
        g = shard_tensor(p + "mlp.gate_proj.weight", sd[p + "mlp.gate_proj.weight"], rank, world)
        u = shard_tensor(p + "mlp.up_proj.weight", sd[p + "mlp.up_proj.weight"], rank, world)
        d = shard_tensor(p + "mlp.down_proj.weight", sd[p + "mlp.down_proj.weight"], rank, world)
        part = (torch.nn.functional.silu(x @ g.T) * (x @ u.T)) @ d.T
        dist.all_reduce(part)
        ref = layer.mlp(x)
        ok = ok and torch.allclose(part, ref, atol=1e-5)
        # 3. vocab-parallel argmax with (value, lowest index) merge == full argmax
        h = torch.randn(7, model.config.hidden_size, generator=torch.Generator().manual_seed(4))
        W = sd["lm_head.weight"]
        Wl = shard_tensor("lm_head.weight", W, rank, world)
        logits = h @ Wl.T
        val, idx = logits.max(-1)
        idx = idx + rank * Wl.shape[0]
        vals = [torch.empty_like(val) for _ in range(world)]
        idxs = [torch.empty_like(idx) for _ in range(world)]
        dist.all_gather(vals, val); dist.all_gather(idxs, idx)
        best_v, best_i = vals[0].clone(), idxs[0].clone()
        for v2, i2 in zip(vals[1:], idxs[1:]):
            take = (v2 > best_v) | ((v2 == best_v) & (i2 < best_i))
            best_v = torch.where(take, v2, best_v); best_i = torch.where(take, i2, best_i)
        ok = ok and torch.equal(best_i, (h @ W.T).argmax(-1))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shards_reassemble_and_block_matches_over_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(0) is True and ret.get(1) is True


def test_divisibility_rules():
    from oracle.hf_model import make_config
    cfg = make_config("qwen2.5-coder-7b")
    check_divisibility(cfg, 2)
    check_divisibility(cfg, 4)
    with pytest.raises((ValueError, NotImplementedError)):
        check_divisibility(cfg, 8)          # 28 Q / 4 KV heads: the reference asserts here too (qwen3.py:31,34)
    with pytest.raises(ValueError):
        check_divisibility(make_config("tiny"), 4)    # 2 kv heads
