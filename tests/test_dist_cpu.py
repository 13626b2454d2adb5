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
