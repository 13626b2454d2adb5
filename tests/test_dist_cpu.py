"""CPU (gloo, world_size 2 and 3): the partition + halo index plan of hidenn-fem_b200/dist.py.
Local arithmetic here is the oracle (numpy closed form); the test itself does the pack/unpack indexing the
CUDA pack kernels do on the GPU, so no product compute path runs on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import relmax
from oracle import closed_form as cf


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _local_oracle(m):
    fmask, umask = ~m.boundary_mask, ~m.dirichlet_mask
    rng_u = 1e-3 * (2.0 * _hash(m.global_node_id) - 1.0)
    U = np.where(umask[:, None], rng_u, 0.0)
    xg, wg = cf.triangle_gauss_points(4)
    xi1, w1 = cf.interval_gauss_points(2)
    loss, dX, dU = cf.tri_energy_full(m.node_coords, U, m.connectivity, cf.plane_stress_C(), xg, wg, None, m.neumann_edges, xi1, w1)
    return loss, dX[fmask], dU[umask], fmask, umask


def _hash(gid):
    from hidenn_fem_b200 import meshgen
    return np.stack([meshgen._hash_u01(gid * 2 + (1 << 50), 7), meshgen._hash_u01(gid * 2 + 1 + (1 << 50), 7)], 1)


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hidenn_fem_b200 import meshgen, dist as hd
        nx, ny = 41, 21
        glob = meshgen.plate_mesh(nx, ny, jitter=0.25, diag="random", seed=0, ordering="morton")
        if mode == "strip":
            m = hd.strip_mesh(nx, ny, rank, world, jitter=0.25, diag="random", seed=0, ordering="morton")
            cand = hd.strip_candidates(m)
        else:
            m = hd.partition_elements(glob, world, rank)
            cand = m.global_node_id
        shared = hd.shared_ids_from_candidates(hd.gather_candidates(cand))
        loss, gx, gu, fmask, umask = _local_oracle(m)
        plan = hd.build_halo_plan(m.global_node_id, fmask, umask, shared)
        S = shared.shape[0]
        buf = torch.zeros(plan.buffer_len, dtype=torch.float64)
        buf[0] = float(loss)
        bx, bu = buf[2:2 + 2 * S].view(S, 2), buf[2 + 2 * S:].view(S, 2)
        bx[torch.from_numpy(plan.x_pos)] = torch.from_numpy(gx[plan.x_rows])
        bu[torch.from_numpy(plan.u_pos)] = torch.from_numpy(gu[plan.u_rows])
        dist.all_reduce(buf)
        gx[plan.x_rows] = bx[torch.from_numpy(plan.x_pos)].numpy()
        gu[plan.u_rows] = bu[torch.from_numpy(plan.u_pos)].numpy()
        # reference: the global mesh on one "device"
        gl, ggx, ggu, gf, gum = _local_oracle(glob)
        pos = {g: i for i, g in enumerate(glob.global_node_id)}
        loc = np.array([pos[g] for g in m.global_node_id])
        full_gx = np.zeros((glob.node_coords.shape[0], 2)); full_gx[gf] = ggx
        full_gu = np.zeros((glob.node_coords.shape[0], 2)); full_gu[gum] = ggu
        ex = relmax(gx, full_gx[loc][fmask])
        eu = relmax(gu, full_gu[loc][umask])
        el = abs(float(buf[0]) - gl) / abs(gl)
        ne = torch.tensor([m.connectivity.shape[0]], dtype=torch.int64)
        dist.all_reduce(ne)
        q.put((rank, el, ex, eu, int(ne.item()), glob.connectivity.shape[0], S))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "strip"), (3, "strip"), (2, "blocks")])
def test_partition_and_halo_allreduce(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, el, ex, eu, ne_sum, ne_glob, S in res:
        assert ne_sum == ne_glob, "element blocks must partition the mesh"
        assert S > 0
        assert el < 1e-12 and ex < 1e-12 and eu < 1e-12, (rank, el, ex, eu)


def test_halo_plan_indices():
    from hidenn_fem_b200 import dist as hd
    gid = np.array([10, 3, 7, 42, 5])
    free = np.array([True, False, True, True, True])
    ufree = np.array([True, True, False, True, True])
    shared = np.array([3, 5, 7, 99])
    p = hd.build_halo_plan(gid, free, ufree, shared)
    assert list(p.local_pos) == [0, 1, 2] and list(p.local_node) == [1, 4, 2]
    # node 1 (gid 3) has a fixed coordinate: no x row; node 2 (gid 7) has a fixed displacement: no u row
    assert list(p.x_pos) == [1, 2] and list(p.x_rows) == [3, 1]
    assert list(p.u_pos) == [0, 1] and list(p.u_rows) == [1, 3]
    assert p.buffer_len == 2 + 4 * 4
    assert list(hd.shared_ids_from_candidates([np.array([1, 2, 3]), np.array([3, 4]), np.array([4, 4, 9])])) == [3, 4]


def _lbfgs_worker(rank, world, port, q, vector_free=True):
    """ShardedLBFGS on the real strip partition: the local energy / gradients come from the numpy oracle, the halo
    sums from the same index plan the CUDA pack kernels use; compared with torch.optim.LBFGS on the global mesh."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hidenn_fem_b200 import meshgen, dist as hd
        from hidenn_fem_b200.optim import ShardedLBFGS
        nx, ny = 21, 11
        kw = dict(jitter=0.2, diag="random", seed=0, ordering="morton")
        xg, wg = cf.triangle_gauss_points(4)
        xi1, w1 = cf.interval_gauss_points(2)
        Cm = cf.plane_stress_C(E=1.0, nu=0.3)

        class OracleLoss(torch.autograd.Function):          # numpy closed form behind autograd (CPU test only)
            @staticmethod
            def forward(ctx, u_free, m, umask, exchange):
                U = np.zeros((m.node_coords.shape[0], 2)); U[umask] = u_free.detach().numpy()
                t_q = np.tile(np.array([1e-3, 0.0]), (m.neumann_edges.shape[0], xi1.shape[0], 1))
                loss, dX, dU = cf.tri_energy_full(m.node_coords, U, m.connectivity, Cm, xg, wg, None, m.neumann_edges, xi1, w1,
                                                  t_q=t_q)
                loss, gu = exchange(float(loss), dU[umask])
                ctx.gu = torch.from_numpy(gu)
                return torch.tensor(loss, dtype=torch.float64)

            @staticmethod
            def backward(ctx, go):
                return ctx.gu * go, None, None, None

        def train(m, exchange, weights, group_on):
            umask = ~m.dirichlet_mask
            u = torch.nn.Parameter(torch.zeros(int(umask.sum()), 2, dtype=torch.float64))
            opt = (ShardedLBFGS([u], max_iter=8, history_size=5, weights=weights, vector_free=vector_free) if group_on
                   else torch.optim.LBFGS([u], max_iter=8, history_size=5))
            losses = []

            def closure():
                opt.zero_grad()
                l = OracleLoss.apply(u, m, umask, exchange)
                l.backward()
                return l
            for _ in range(3):
                losses.append(float(opt.step(closure)))
            return losses, u.detach().numpy(), umask

        m = hd.strip_mesh(nx, ny, rank, world, **kw)
        cands = hd.gather_candidates(hd.strip_candidates(m))
        shared = hd.shared_ids_from_candidates(cands)
        plan = hd.build_halo_plan(m.global_node_id, ~m.boundary_mask, ~m.dirichlet_mask, shared)
        S = shared.shape[0]
        _, wu = hd.owner_weights(plan, hd.shared_owner_from_candidates(cands, shared), rank, int((~m.boundary_mask).sum()),
                                 int((~m.dirichlet_mask).sum()))

        def exchange(loss, gu):
            buf = torch.zeros(1 + 2 * S, dtype=torch.float64)
            buf[0] = loss
            bu = buf[1:].view(S, 2)
            bu[torch.from_numpy(plan.u_pos)] = torch.from_numpy(gu[plan.u_rows])
            dist.all_reduce(buf)
            gu = gu.copy()
            gu[plan.u_rows] = bu[torch.from_numpy(plan.u_pos)].numpy()
            return float(buf[0]), gu

        ls, u_loc, umask = train(m, exchange, [torch.from_numpy(wu)], True)
        glob = meshgen.plate_mesh(nx, ny, **kw)
        lg, u_glob, gum = train(glob, lambda l, g: (l, g), None, False)
        pos = {g: i for i, g in enumerate(glob.global_node_id)}
        loc = np.array([pos[g] for g in m.global_node_id])
        full = np.zeros((glob.node_coords.shape[0], 2)); full[gum] = u_glob
        q.put((rank, relmax(np.array(ls), np.array(lg)), relmax(u_loc, full[loc][umask]), float(wu.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("vector_free", [True, False])
def test_sharded_lbfgs_follows_global_lbfgs(vector_free):
    """vector_free=True: one batched all-reduce of inner products per iteration (Gram-matrix two-loop recursion);
    False: the textbook recursion with one all-reduce per inner product."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lbfgs_worker, args=(r, world, port, q, vector_free)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, el, eu, nw in res:
        assert el < 1e-9 and eu < 1e-7, (rank, el, eu)
    # every free displacement row is owned exactly once across ranks
    from hidenn_fem_b200 import meshgen
    glob = meshgen.plate_mesh(21, 11, jitter=0.2, diag="random", seed=0, ordering="morton")
    assert sum(r[3] for r in res) == float((~glob.dirichlet_mask).sum())


def test_sharded_lbfgs_equals_torch_lbfgs_single_process():
    from hidenn_fem_b200.optim import ShardedLBFGS
    torch.manual_seed(0)
    n = 120
    A = torch.randn(n, n, dtype=torch.float64)
    A = A @ A.T / n + torch.eye(n, dtype=torch.float64)
    b = torch.randn(n, dtype=torch.float64)

    def run(cls, **kw):
        x = torch.nn.Parameter(torch.zeros(n, dtype=torch.float64))
        y = torch.nn.Parameter(torch.ones(5, 2, dtype=torch.float64))
        opt = cls([x, y], **kw)

        def closure():
            opt.zero_grad()
            l = 0.5 * x @ A @ x - b @ x + (y ** 4).sum() + y.sum() * x[:3].sum()
            l.backward()
            return l
        ls = [float(opt.step(closure).detach()) for _ in range(5)]
        return np.array(ls), x.detach().numpy(), y.detach().numpy()

    for kw in (dict(), dict(lr=0.5, max_iter=7, history_size=3), dict(lr=0.3, max_iter=30, history_size=5),
               dict(lr=0.2, max_iter=40, history_size=20)):      # > 8 stored pairs: the history storage grows
        l1, x1, y1 = run(torch.optim.LBFGS, **kw)
        for vf in (False, True):
            l2, x2, y2 = run(ShardedLBFGS, vector_free=vf, **kw)
            assert relmax(l2, l1) < 1e-10 and relmax(x2, x1) < 1e-7 and relmax(y2, y1) < 1e-7, (kw, vf)


def test_partition_elements_keeps_unsorted_neumann_edges():
    """Edges stored with first id > second id (gmsh output, meshes renumbered by reorder_for_locality) must not be
    dropped by the partitioner; the stored orientation is kept (the reference's edge rule depends on it, Q3)."""
    from hidenn_fem_b200 import meshgen, dist as hd
    glob = meshgen.plate_mesh(33, 17, jitter=0.25, diag="random", seed=0, ordering="random")
    ed = glob.neumann_edges.copy()
    ed[::2] = ed[::2, ::-1]                               # every other edge descending
    glob.neumann_edges = ed
    assert (ed[:, 0] > ed[:, 1]).any() and (ed[:, 0] < ed[:, 1]).any()
    world, seen = 3, []
    for rank in range(world):
        m = hd.partition_elements(glob, world, rank)
        gid = m.global_node_id[m.neumann_edges]              # back to the generating grid's ids
        seen.append(gid)
    got = np.concatenate(seen)
    want = glob.global_node_id[ed]
    assert got.shape[0] == ed.shape[0]                       # every edge on exactly one rank
    key = lambda a: a[:, 0] * (1 << 32) + a[:, 1]            # orientation-sensitive
    assert np.array_equal(np.sort(key(got)), np.sort(key(want)))


@pytest.mark.parametrize("world,mode", [(2, "strip"), (4, "strip"), (3, "blocks")])
def test_peer_tables_emulated_exchange(world, mode):
    """Index tables of the peer-memory halo exchange (dist.build_peer_tables), replayed with numpy: every rank puts the
    partial gradients of its shared nodes into the other holders' receive blocks (slot k of the pair's common list) and
    completes its own rows by summing all holders in ascending rank order -> the partitioned gradients equal the
    single-mesh oracle, and every holder of a node ends with bit-identical values."""
    from hidenn_fem_b200 import meshgen, dist as hd
    nx, ny = 41, 21
    glob = meshgen.plate_mesh(nx, ny, jitter=0.25, diag="random", seed=0, ordering="morton")
    parts, cands = [], []
    for rank in range(world):
        if mode == "strip":
            m = hd.strip_mesh(nx, ny, rank, world, jitter=0.25, diag="random", seed=0, ordering="morton")
            cands.append(hd.strip_candidates(m))
        else:
            m = hd.partition_elements(glob, world, rank)
            cands.append(m.global_node_id)
        parts.append(m)
    shared = hd.shared_ids_from_candidates(cands)
    loc, tabs = [], []
    for rank, m in enumerate(parts):
        loss, gx, gu, fmask, umask = _local_oracle(m)
        loc.append([loss, gx, gu, fmask, umask])
        tabs.append(hd.build_peer_tables(cands, shared, rank, m.global_node_id, fmask, umask))
    smax = tabs[0].smax
    assert all(t.smax == smax for t in tabs) and smax > 0
    recv = np.zeros((world, world, smax, 4))                      # [receiver, sender, slot, (gx, gu)]
    for r, t in enumerate(tabs):                                   # push
        gx, gu = loc[r][1], loc[r][2]
        for xr, ur, q, k in zip(t.send_xrow, t.send_urow, t.send_peer, t.send_k):
            recv[q, r, k, :2] = gx[xr] if xr >= 0 else 0.0
            recv[q, r, k, 2:] = gu[ur] if ur >= 0 else 0.0
    done = {}
    for r, t in enumerate(tabs):                                   # pull
        gx, gu = loc[r][1].copy(), loc[r][2].copy()
        assert set(t.wait_ranks.tolist()) == set(t.send_peer.tolist())
        for j in range(t.node_xrow.size):
            acc = np.zeros(4)
            ranks = t.src_rank[t.node_off[j]:t.node_off[j + 1]]
            assert (np.diff(ranks) > 0).all() and r in ranks and len(ranks) >= 2
            for s in range(t.node_off[j], t.node_off[j + 1]):
                q = t.src_rank[s]
                if q == r:
                    own = np.concatenate([loc[r][1][t.node_xrow[j]] if t.node_xrow[j] >= 0 else np.zeros(2),
                                          loc[r][2][t.node_urow[j]] if t.node_urow[j] >= 0 else np.zeros(2)])
                    acc = acc + own
                else:
                    acc = acc + recv[r, q, t.src_k[s]]
            if t.node_xrow[j] >= 0:
                gx[t.node_xrow[j]] = acc[:2]
            if t.node_urow[j] >= 0:
                gu[t.node_urow[j]] = acc[2:]
            gid = int(parts[r].global_node_id[t.local_node[j]])
            done.setdefault(gid, []).append(acc)
        loc[r][1], loc[r][2] = gx, gu
    for gid, vals in done.items():                                 # all holders: same bits
        assert all(np.array_equal(v, vals[0]) for v in vals), gid
    gl, ggx, ggu, gf, gum = _local_oracle(glob)
    pos = {g: i for i, g in enumerate(glob.global_node_id)}
    full_gx = np.zeros((glob.node_coords.shape[0], 2)); full_gx[gf] = ggx
    full_gu = np.zeros((glob.node_coords.shape[0], 2)); full_gu[gum] = ggu
    for r, m in enumerate(parts):
        idx = np.array([pos[g] for g in m.global_node_id])
        assert relmax(loc[r][1], full_gx[idx][loc[r][3]]) < 1e-12
        assert relmax(loc[r][2], full_gu[idx][loc[r][4]]) < 1e-12
    assert abs(sum(l[0] for l in loc) - gl) <= 1e-12 * abs(gl)


def test_sharded_lbfgs_strong_wolfe_follows_the_stock_optimiser():
    """ShardedLBFGS(line_search_fn="strong_wolfe") on one process: same trajectory as torch.optim.LBFGS with its
    strong-Wolfe line search (both the two-loop and the vector-free form); FP32 history stays within single precision."""
    from hidenn_fem_b200.optim import ShardedLBFGS
    torch.manual_seed(0)
    A = torch.randn(40, 40, dtype=torch.float64)
    A = A @ A.T + torch.eye(40, dtype=torch.float64)
    b = torch.randn(40, dtype=torch.float64)
    f = lambda x: 0.5 * x @ A @ x - b @ x + 0.1 * (x ** 4).sum()

    def run(make):
        x = torch.nn.Parameter(torch.zeros(40, dtype=torch.float64))
        o = make([x])
        tr = []
        for _ in range(4):
            def c():
                o.zero_grad()
                l = f(x)
                l.backward()
                return l
            tr.append(float(o.step(c).detach()))
        return tr, x.detach().clone()
    kw = dict(max_iter=8, history_size=10, line_search_fn="strong_wolfe")
    ref_tr, ref_x = run(lambda p: torch.optim.LBFGS(p, **kw))
    assert ref_tr[-1] < ref_tr[1] < ref_tr[0] + 1e-12
    for extra, tol in ((dict(), 1e-12), (dict(vector_free=True), 1e-12), (dict(history_dtype=torch.float32), 1e-6)):
        tr, x = run(lambda p: ShardedLBFGS(p, **kw, **extra))
        assert np.allclose(tr[1:], ref_tr[1:], rtol=tol, atol=0), (extra, tr, ref_tr)
        assert float((x - ref_x).abs().max()) < 1e3 * tol
    with pytest.raises(RuntimeError):
        ShardedLBFGS([torch.nn.Parameter(torch.zeros(2))], line_search_fn="armijo")
