"""world_size-2 (or more) worker for the CPU test of the multi-GPU halo logic (gloo backend).

Each rank builds its halo plan with the product's host code (pamg_halo_plan), fills the strips of its own
parents from a seeded field with numpy, exchanges the strips of the cut faces with its peers exactly as
exchange_halo() in pamg_api.cu does (send slot range -> peer, receive into the contiguous strip range),
and checks every local strip against the whole-mesh oracle update_overlaps.  No GPU, no product compute."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_api as orc  # noqa: E402
from pamg_pkg import pamg  # noqa: E402


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    kp, n = 2, 3
    S = 2 ** n
    mesh = pamg.Mesh.synthetic(kp, world)
    per = 4 ** kp
    pf = np.arange(world + 1, dtype=np.int32) * per
    rng = np.random.Generator(np.random.MT19937(99))
    T = rng.random((mesh.U, 4 ** n, 3))
    # oracle on the whole mesh
    p = orc.intended_params(n, 1)
    s = orc.Semi(p, mesh.X, mesh.neig, mesh.fneig, mesh.dir)
    s.field(orc.TNEW)[:] = T
    s.update_overlaps(1)
    ref = s.overlap(1)
    plan = pamg.halo_plan(mesh, 1, world, pf, rank)
    first, UL = plan["first"], plan["U_local"]
    surf = np.zeros(3 * S, np.int32)
    orc.lib().orc_surf_ele(n, surf)
    surf = surf.reshape(3, S)
    space = np.zeros((plan["nstrips"] + plan["nsend"], S, 3))
    for u in range(UL):
        for mf in range(3):
            lf = u * 3 + mf
            d = plan["dst_strip"][lf]
            if d < 0:
                continue
            for i in range(S):
                slot = S - 1 - i if plan["rev"][lf] else i
                space[d, slot] = T[first + u, surf[mf, i] - 1]
    reqs, bufs = [], []
    for part, nfaces, sb, se in plan["peers"]:
        send = torch.from_numpy(space[plan["nstrips"] + se: plan["nstrips"] + se + nfaces].copy())
        recv = torch.zeros_like(send)
        reqs.append(dist.isend(send, int(part)))
        reqs.append(dist.irecv(recv, int(part)))
        bufs.append((sb, nfaces, recv))
    for r in reqs:
        r.wait()
    for sb, nfaces, recv in bufs:
        space[sb: sb + nfaces] = recv.numpy()
    got = space[plan["strip_of"]].reshape(UL, 3, S, 3)
    loc = ref[first: first + UL]
    interior = mesh.neig[first: first + UL] != 0
    ok = bool(np.array_equal(got[interior], loc[interior]))
    ncut = int(sum(p_[1] for p_ in plan["peers"]))
    flags = [None] * world
    dist.all_gather_object(flags, (ok, ncut))
    if rank == 0:
        print("DIST_OK" if all(f[0] for f in flags) and sum(f[1] for f in flags) > 0 else "DIST_FAIL", flags)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
