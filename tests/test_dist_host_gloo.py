"""N > 1 host logic on CPU: two processes over gloo (no GPU).  Each rank owns the row stripe the reference's
Spmv::preprocess would give pipe `rank` (Spmv.cpp:334-364), derives the x windows its 1024-row slices stage, asks the
library's host-side planner (cask_b200_halo_plan_host - the routine the GPU path runs) which column ranges to fetch
from which peer, swaps the requests, moves exactly those entries over gloo into a full-layout x whose foreign part
is otherwise poisoned with NaN, and multiplies its stripe with the CPU oracle.  The result must equal the global
oracle product bit for bit: nothing a row needs may be missing from the plan."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import cask_b200 as cb
from oracle import oraclebind as O
O.build()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
out = {}
for name, gen, N in (("poisson2d", O.gen_poisson2d, 75), ("poisson3d27", O.gen_poisson3d27, 17), ("convdiff3d7", O.gen_convdiff3d7, 19)):
    n, rp, ci, va = gen(N)                       # odd sizes: n %% world != 0, the last stripe takes the remainder
    r0, nr = cb.shard_rows(n, world, rank)
    assert (r0, nr) == ((n // world) * rank, n // world if rank < world - 1 else n - (n // world) * (world - 1))
    # x windows of every 1024-row slice of the stripe: runs of consecutive 16-column granules
    c0s, lens = [], []
    for s0 in range(r0, r0 + nr, 1024):
        cols = ci[rp[s0]:rp[min(s0 + 1024, r0 + nr)]]
        g = np.unique(cols // 16)
        if len(g) == 0:
            continue
        brk = np.flatnonzero(np.diff(g) > 1)
        starts = np.concatenate(([g[0]], g[brk + 1])); ends = np.concatenate((g[brk], [g[-1]])) + 1
        c0s += list(starts * 16); lens += list(np.minimum(ends * 16, n) - starts * 16)
    recv = cb.halo_plan_host(n, world, rank, c0s, lens)
    # every requested range lies inside its owner's stripe and outside mine
    for q, c0, ln in recv:
        q0, qn = cb.shard_rows(n, world, q)
        assert q != rank and q0 <= c0 and c0 + ln <= q0 + qn and ln > 0
    # swap the requests: what each peer wants from me
    allreq = [None] * world
    dist.all_gather_object(allreq, recv)
    send = [(dst, c0, ln) for dst in range(world) for (q, c0, ln) in allreq[dst] if q == rank]
    x = np.random.default_rng(11).random(n)
    xf = torch.full((n,), float("nan"), dtype=torch.float64)
    xf[r0:r0 + nr] = torch.from_numpy(x[r0:r0 + nr])
    ops = []
    for dst, c0, ln in send:
        ops.append(dist.P2POp(dist.isend, xf[c0:c0 + ln].clone(), dst))
    bufs = []
    for q, c0, ln in recv:
        b = torch.empty(ln, dtype=torch.float64); bufs.append((c0, ln, b))
        ops.append(dist.P2POp(dist.irecv, b, q))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for c0, ln, b in bufs:
        xf[c0:c0 + ln] = b
    # stripe product with the oracle on the haloed vector (sliceRows semantics: row_ptr rebased, columns global)
    lrp = (rp[r0:r0 + nr + 1] - rp[r0]).astype(np.int32)
    lci = ci[rp[r0]:rp[r0 + nr]]; lva = va[rp[r0]:rp[r0 + nr]]
    needed = np.unique(lci)
    xs = xf.numpy()
    assert not np.isnan(xs[needed]).any(), "a column the stripe references was not delivered"
    xs = np.nan_to_num(xs, nan=0.0)
    full_rp = np.concatenate((lrp, np.full(n - nr, lrp[-1], np.int32)))   # pad to a square system for the oracle
    got = O.csr_dot(n, full_rp, lci, lva, xs)[:nr]
    exp = O.csr_dot(n, rp, ci, va, x)[r0:r0 + nr]
    halo_doubles = sum(ln for _, _, ln in recv)
    # a local dot product all-reduced over gloo: the solver's p.Ap / r.r step
    part = torch.tensor([float(np.dot(got, got))], dtype=torch.float64)
    dist.all_reduce(part)
    out[name] = {"bitexact": bool(np.array_equal(got, exp)), "halo": halo_doubles, "npeers": len({q for q, _, _ in recv}),
                 "yy": float(part.item()), "yy_exp": float(np.dot(O.csr_dot(n, rp, ci, va, x), O.csr_dot(n, rp, ci, va, x)))}
res = [None] * world
dist.all_gather_object(res, out)
if rank == 0:
    print("RESULT " + json.dumps(res))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_moves_exactly_what_the_stripe_needs(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-6000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert len(res) == world
    for rank, r in enumerate(res):
        for name, v in r.items():
            assert v["bitexact"], (rank, name, v)
            assert v["halo"] > 0 and 1 <= v["npeers"] <= 2, (rank, name, v)   # stencils: neighbours only
            assert abs(v["yy"] - v["yy_exp"]) <= 1e-12 * abs(v["yy_exp"]), (rank, name, v)
    # interior ranks talk to two neighbours, the ends to one
    if world == 3:
        assert res[1]["poisson2d"]["npeers"] == 2 and res[0]["poisson2d"]["npeers"] == 1


def test_reference_arm_under_torchrun_prints_once(tmp_path):
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone prints the line."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29711", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--no-cpu"]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["impl"] == "reference" and json.loads(lines[0])["n_gpus"] == 2


SPARSE_WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import cask_b200 as cb
from oracle import oraclebind as O
O.build()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n, rp, ci, va = O.gen_rmat(12, 8, 3)
# stripes of equal nonzero count (rank 0 owns the few, long hub rows), as the R-MAT leg of bench.py cuts them
bounds = [0] + [int(np.searchsorted(rp, rp[-1] * q // world)) for q in range(1, world)] + [n]
r0, nr = bounds[rank], bounds[rank + 1] - bounds[rank]
lci = ci[rp[r0]:rp[r0 + nr]]
need, counts = np.unique(lci, return_counts=True)          # the compact numbering: referenced columns, ascending ...
need = need.astype(np.int32)
HUB = os.environ.get("SPARSE_HUB") == "1"
if HUB:                                                    # ... or grouped by owner, every group's most referenced columns first
    owner = np.searchsorted(np.asarray(bounds[1:]), need, side="right")
    need = need[np.lexsort((need, -counts, owner))]
seg = cb.sparse_segments_host(bounds, need)
assert seg[0] == 0 and seg[-1] == len(need) and all(seg[q] <= seg[q + 1] for q in range(world))
for q in range(world):                                     # segment q holds exactly the columns rank q owns
    s = need[seg[q]:seg[q + 1]]
    assert len(s) == 0 or (bounds[q] <= s.min() and s.max() < bounds[q + 1])
segs = [torch.zeros(world + 1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(segs, torch.from_numpy(seg.copy()))
all_seg = np.stack([t.numpy() for t in segs])
send_off, dst_off = cb.sparse_send_plan_host(all_seg, rank)
# request lists: what every rank needs (the GPU path all-gathers the padded lists and cuts its pieces)
lists = [None] * world
dist.all_gather_object(lists, need)
x = np.random.default_rng(3).standard_normal(n)
xp = torch.full((max(len(need), 1),), float("nan"), dtype=torch.float64)     # the compact x, poisoned
xp[seg[rank]:seg[rank + 1]] = torch.from_numpy(x[need[seg[rank]:seg[rank + 1]]])   # own segment, filled locally
ops, keep = [], []
for q in range(world):
    if q == rank:
        continue
    want = lists[q][all_seg[q][rank]:all_seg[q][rank + 1]]      # columns of MINE that rank q references
    assert len(want) == send_off[q + 1] - send_off[q] and dst_off[q] == all_seg[q][rank]
    assert len(want) == 0 or (bounds[rank] <= want.min() and want.max() < bounds[rank + 1])
    if len(want):
        buf = torch.from_numpy(x[want].copy()); keep.append(buf)  # a rank only ever reads its OWN slice of x
        ops.append(dist.P2POp(dist.isend, buf, q))
    if seg[q + 1] > seg[q]:
        ops.append(dist.P2POp(dist.irecv, xp[seg[q]:seg[q + 1]], q))
if ops:
    for r in dist.batch_isend_irecv(ops):
        r.wait()
got = xp.numpy()[:len(need)]
assert not np.isnan(got).any() and np.array_equal(got, x[need])                # every referenced entry arrived, in place
# the relabelled stripe times the compact x == the global product
inv = np.full(n, -1, np.int64); inv[need] = np.arange(len(need))
relabel = inv[lci].astype(np.int32)
assert (relabel >= 0).all()
lrp = (rp[r0:r0 + nr + 1] - rp[r0]).astype(np.int32)
y = O.csr_dot(nr, lrp, relabel, va[rp[r0]:rp[r0 + nr]], got) if nr else np.zeros(0)
exp = O.csr_dot(n, rp, ci, va, x)[r0:r0 + nr]
# the oracle's dot (CsrMatrix::dot) sums a row in ascending COLUMN order: identical bits in column order, a permutation
# of the same products (tolerance of the gather kernels, relative to sum |a_ij x_j|) when the hubs come first
scale = O.csr_dot(n, rp, ci, np.abs(va), np.abs(x))[r0:r0 + nr]
ok = bool(np.array_equal(y, exp)) if not HUB else bool((np.abs(y - exp) <= 1e-12 * np.maximum(scale, 1e-300)).all())
t = torch.tensor([1.0 if ok else 0.0, float(len(need)), float(send_off[-1])], dtype=torch.float64)
tl = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(tl, t)
if rank == 0:
    print("RESULT " + json.dumps({"ok": [bool(v[0]) for v in tl], "need": [int(v[1]) for v in tl], "sent": [int(v[2]) for v in tl], "n": n}))
dist.destroy_process_group()
'''


@pytest.mark.parametrize("hub", [0, 1])
@pytest.mark.parametrize("world", [2, 3])
def test_sparse_exchange_plan_over_gloo(world, hub, tmp_path):
    """The sparse exchange of a row-sharded gather plan, host arithmetic only (cask_b200_sparse_segments_host /
    cask_b200_sparse_send_plan_host - the routines dist.cu runs), with gloo standing in for NVLink: every rank receives
    exactly the x entries its rows reference into its compactly renumbered x, nothing else travels, and the relabelled
    stripe times the compact x equals the global product bit for bit.  hub = 1: the compact numbering is grouped by owner
    with every group's most referenced columns first, as dist.cu builds it."""
    script = tmp_path / "worker.py"
    script.write_text(SPARSE_WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29640 + 4 * hub + world), str(script)]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, SPARSE_HUB=str(hub)))
    assert p.returncode == 0, p.stdout[-4000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert all(res["ok"]) and len(res["ok"]) == world
    assert all(0 < k < res["n"] for k in res["need"])          # every rank references a strict subset of the columns
    assert 0 < sum(res["sent"]) < sum(res["need"])          # what travels = what is referenced minus the own segments
