#!/usr/bin/env python
"""BASELINE.json config 5 across GPUs: 1 B documents (Zipf string keys, 20 % MISSING/NULL), range-partitioned over the
ranks (STRONG scaling: the total is fixed), columns resident in HBM.  Run under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/scale_config5.py

A step = this rank's scan + the Intermediate merge (the group table is direct-indexed, so the merge is an NCCL
in-place NCCL all-reduce of the accumulator words, one call per run of sum / min / max words, stream-ordered); timed with CUDA events, max over ranks.
The finalisation on the host (ComputeFinal of 100 002 groups) is timed separately.  Rank 0 checks the merged result
against torch reductions all-reduced over the ranks' tensors, and prints one JSON line.
FS_SCALE shrinks the row count; FS_WEAK=1 keeps 1 B x FS_SCALE rows PER RANK instead (weak scaling)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import query_b200 as q  # noqa: E402
from query_b200 import dist as qd  # noqa: E402

C_MISSING, C_NULL, C_INT, C_STRING = 0, 1, 4, 6


def tags_of(r, present):
    tg = torch.full(r.shape, present, dtype=torch.uint8, device=r.device)
    tg[r == 0] = C_MISSING
    tg[r == 1] = C_NULL
    return tg


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    q.init(local)
    scale = float(os.environ.get("FS_SCALE", "1"))
    weak = os.environ.get("FS_WEAK", "0") == "1"
    total = int(1_000_000_000 * scale) * (world if weak else 1)
    lo, hi = qd.row_range(total, rank, world)
    n = hi - lo
    vocab = 100_000
    words = sorted("w%06d-%x" % (i, (i * 2654435761) & 0xffffff) for i in range(vocab))
    common = torch.Generator(device=dev).manual_seed(4)       # same on every rank: popularity rank -> dictionary code
    perm = torch.randperm(vocab, generator=common, device=dev).int()
    g = torch.Generator(device=dev).manual_seed(1000 + rank)   # this rank's documents
    w = torch.arange(1, vocab + 1, dtype=torch.float64, device=dev).pow(-1.1)
    cdf = torch.cumsum(w / w.sum(), 0)
    code = torch.empty(n, dtype=torch.int32, device=dev)
    ktag = torch.empty(n, dtype=torch.uint8, device=dev)
    v = torch.empty(n, dtype=torch.int64, device=dev)
    vtag = torch.empty(n, dtype=torch.uint8, device=dev)
    chunk = 50_000_000
    for a in range(0, n, chunk):
        m = min(chunk, n - a)
        u = torch.rand(m, generator=g, device=dev, dtype=torch.float64)
        code[a:a + m] = perm[torch.searchsorted(cdf, u).clamp_(max=vocab - 1)]
        ktag[a:a + m] = tags_of(torch.randint(0, 10, (m,), generator=g, device=dev), C_STRING)
        v[a:a + m] = torch.randint(-1000, 1_000_000, (m,), generator=g, device=dev, dtype=torch.int64)
        vtag[a:a + m] = tags_of(torch.randint(0, 10, (m,), generator=g, device=dev), C_INT)
        del u
    t = q.Table(["k", "v"])
    t.set_column_device("k", code, tags=ktag, dictionary=words)
    t.set_column_device("v", v, tags=vtag)
    qd.agree_dictionaries_and_stats(t)
    t.seal()
    qq = q.Query(t, "d", "((`d`.`v`) is not missing)", ["(`d`.`k`)"],
                 ["count(*)", "count((`d`.`v`))", "sum((`d`.`v`))", "min((`d`.`v`))", "max((`d`.`v`))"])
    qq.set_timing(False)
    qq.set_stream(torch.cuda.current_stream().cuda_stream)  # the events below are recorded on this stream
    dq = qd.DistributedQuery(qq)
    res = dq.execute()
    res = dq.execute()
    K = int(os.environ.get("FS_STEPS", "5"))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_ms, fin_ms = [], []
    for _ in range(K):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        dq.launch()
        ev1.record()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = dq.collect()
        ng = res.num_groups
        fin_ms.append((time.perf_counter() - t0) * 1e3)
        dev_ms.append(ev0.elapsed_time(ev1))
    ms = sorted(dev_ms)[len(dev_ms) // 2]
    fms = sorted(fin_ms)[len(fin_ms) // 2]
    if world > 1:
        tm = torch.tensor([ms, fms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms, fms = float(tm[0]), float(tm[1])
    # ---- check (every rank reduces its own tensors; the reference is all-reduced) ------------------------------
    passing = vtag != C_MISSING
    gid = torch.where(ktag == C_STRING, code.long() + 2, ktag.long())[passing]
    isint = (vtag == C_INT)[passing]
    vv = v[passing]
    G = vocab + 2
    cnt = torch.bincount(gid, minlength=G)
    cntv = torch.bincount(gid[isint], minlength=G)
    sm = torch.zeros(G, dtype=torch.int64, device=dev).index_add_(0, gid[isint], vv[isint])
    mn = torch.full((G,), 2 ** 62, dtype=torch.int64, device=dev).scatter_reduce_(0, gid[isint], vv[isint], "amin")
    mx = torch.full((G,), -2 ** 62, dtype=torch.int64, device=dev).scatter_reduce_(0, gid[isint], vv[isint], "amax")
    if world > 1:
        for x, op in ((cnt, dist.ReduceOp.SUM), (cntv, dist.ReduceOp.SUM), (sm, dist.ReduceOp.SUM), (mn, dist.ReduceOp.MIN), (mx, dist.ReduceOp.MAX)):
            dist.all_reduce(x, op=op)
    if rank == 0:
        cnt, cntv, sm, mn, mx = (a.cpu().numpy() for a in (cnt, cntv, sm, mn, mx))
        index = {wd: i + 2 for i, wd in enumerate(words)}
        rows = res.rows()
        assert len(rows) == int((cnt > 0).sum()), (len(rows), int((cnt > 0).sum()))
        for keys, a in rows:
            k = keys[0]
            i = 0 if k is q.MISSING else (1 if k is None else index[k])
            assert a[0] == cnt[i] and a[1] == cntv[i], (k, a, cnt[i], cntv[i])
            if cntv[i]:
                assert float(a[2]) == float(sm[i]) and a[3] == mn[i] and a[4] == mx[i], (k, a, sm[i], mn[i], mx[i])
        info = qq.info
        print(json.dumps({"config": "config5", "n_gpus": world, "scaling": "weak" if weak else "strong", "total_rows": total,
                          "rows_per_gpu": n, "mode": info["mode"], "scan_plus_merge_ms": ms, "rows_per_s": total / (ms * 1e-3),
                          "gb_per_s": info["scan_bytes_per_row"] * total / (ms * 1e6), "finalize_host_ms": fms, "groups": ng,
                          "merge": "none" if world == 1 else "in-place NCCL all-reduce of the direct-indexed table, one call per run of sum / min / max words, stream-ordered",
                          "check": "every group exact vs all-reduced torch reductions"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
