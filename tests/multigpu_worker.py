"""TEST INFRASTRUCTURE -- one rank of the multi-GPU parity test (launched by tests/test_multigpu.py through
torch.distributed.run, one process per GPU, NCCL).  Every rank also computes the WHOLE job alone on its own GPU
(single-rank pipeline) and the distributed results -- run(), run_cyclic(), run_from_host(), peer modes "dma" and
"stores", several consecutive steps so that buffers and barrier channels are reused -- must reproduce it."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynaphopy_b200 import ops as dops, pipeline  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ops = dops.Ops()
    rng = np.random.default_rng(4242)
    dims = np.array([4, 4, 2])
    nc = int(np.prod(dims))
    cell = np.diag([4.0, 4.0, 4.0])
    unit_pos = np.array([[0.0, 0.0, 0.0], [2.0, 2.0, 2.0]])
    grid = np.array([[x, y, z] for z in range(dims[2]) for y in range(dims[1]) for x in range(dims[0])], float) @ cell
    pos = (unit_pos[:, None, :] + grid[None]).reshape(-1, 3)
    typ = np.repeat(np.arange(2), nc).astype(np.int32)
    mass = np.repeat([28.0855, 69.723], nc)
    m = np.array([[mx, my, mz] for mz in range(dims[2]) for my in range(dims[1]) for mx in range(dims[0])])
    q_cart = (m / dims) @ (2.0 * np.pi * np.linalg.inv(cell).T)
    Q, M = len(q_cart), 6
    eig = np.stack([np.linalg.qr(rng.normal(size=(M, M)) + 1j * rng.normal(size=(M, M)))[0].T for _ in range(Q)])
    tc, nch = 256, 3
    T = world * nch * tc
    dt, freq = 0.002, np.arange(0, 40.0, 1.0)                   # L = 500: several overlapping windows
    v = rng.normal(size=(T, 2 * nc, 3))
    t = np.arange(T)[:, None, None] * dt
    v += 0.8 * np.cos(2 * np.pi * 7.0 * t + rng.uniform(0, 2 * np.pi, size=(1, 2 * nc, 3)))
    plan = pipeline.CommensuratePlan(cell, dims, pos, mass, typ, 2, q_cart)
    assert plan.usable() and plan.commensurate.all()
    groups = [dist.new_group([r]) for r in range(world)]          # every rank must take part in every new_group call
    solo = groups[rank]
    failures = []
    for alg in (2, 1, 0):
        ref_pipe = pipeline.MeshPipeline(plan, eig, freq, dt, ops=ops, group=solo, algorithm=alg, mem_coefficients=40,
                                         correlation_weight=1)
        ref = ref_pipe.run(torch.as_tensor(v, device=dev), tc).cpu().numpy()          # (F, Q*M), the whole job on one GPU
        pipe = pipeline.MeshPipeline(plan, eig, freq, dt, ops=ops, algorithm=alg, mem_coefficients=40, correlation_weight=1)
        tl = T // world
        v_blocks = torch.as_tensor(v[rank * tl:(rank + 1) * tl], device=dev)
        v_cyc = torch.as_tensor(np.concatenate([v[a:b] for a, b in pipe.cyclic_chunks(T, tc)]), device=dev)
        v_host = torch.as_tensor(v[rank * tl:(rank + 1) * tl]).pin_memory()

        def check(name, ps_local):
            ps_local = ps_local if torch.is_tensor(ps_local) else torch.as_tensor(np.ascontiguousarray(ps_local), device=dev)
            parts = [torch.empty_like(ps_local) for _ in range(world)]
            dist.all_gather(parts, ps_local.contiguous())
            full = torch.cat(parts, dim=1).cpu().numpy()
            err = float(np.abs(full - ref).max() / np.abs(ref).max())
            if not err < 1e-12:
                failures.append((alg, name, err))
            if rank == 0:
                print("alg %d %-28s max rel err %.2e" % (alg, name, err), flush=True)

        for mode in ("dma", "stores"):
            os.environ["DP_PEER_MODE"] = mode
            for step in range(3):                                  # consecutive steps: send slots, receive buffers, barriers reused
                check("run %s step %d" % (mode, step), pipe.run(v_blocks, tc))
            if alg == 2:
                for step in range(2):
                    check("run_cyclic %s step %d" % (mode, step), pipe.run_cyclic(v_cyc, tc))
            for step in range(2):
                check("run_from_host %s step %d" % (mode, step), pipe.run_from_host(v_host, chunk_frames=tc))
        if alg == 2:
            # run_cyclic above exchanged the half-mesh intermediate (when the pairs (q, -q) deal evenly over the ranks);
            # the same step with the contracted series exchanged, and the intermediate through NCCL instead of peer memory
            used = pipe.intermediate_exchange_plan() is not None
            if rank == 0:
                print("run_cyclic exchanged the %s" % ("half-mesh intermediate" if used else "contracted series"), flush=True)
            os.environ["DP_NO_VC_EXCHANGE"] = "1"
            p2 = pipeline.MeshPipeline(plan, eig, freq, dt, ops=ops, algorithm=alg)
            for step in range(2):
                check("run_cyclic series exchange %d" % step, p2.run_cyclic(v_cyc, tc))
            del os.environ["DP_NO_VC_EXCHANGE"]
            os.environ["DP_NO_PEER_EXCHANGE"] = "1"
            p3 = pipeline.MeshPipeline(plan, eig, freq, dt, ops=ops, algorithm=alg)
            check("run_cyclic intermediate, nccl", p3.run_cyclic(v_cyc, tc))
            del os.environ["DP_NO_PEER_EXCHANGE"]
        os.environ["DP_NO_PEER_EXCHANGE"] = "1"                    # the NCCL all-to-all fallback
        fb = pipeline.MeshPipeline(plan, eig, freq, dt, ops=ops, algorithm=alg, mem_coefficients=40, correlation_weight=1)
        check("run nccl all_to_all", fb.run(v_blocks, tc))
        del os.environ["DP_NO_PEER_EXCHANGE"]
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("FAILURES: %s" % failures if failures else "MULTIGPU PARITY OK (world %d)" % world, flush=True)
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
