"""Development aid: projection kernel time with phases skipped (DP_PF_SKIP: 1 loads, 2 transforms, 4 epilogue)."""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from dynaphopy_b200 import ops as dops
    ops = dops.Ops()
    dims = np.array([16, 16, 20]); nc = int(np.prod(dims)); T = 4096
    mm = np.array([[x, y, z] for z in range(20) for y in range(16) for x in range(16)], dtype=np.int32)
    qb = torch.as_tensor(mm, device=ops.be.device)
    v = torch.randn((T, 2 * nc, 3), dtype=torch.float64, device=ops.be.device)
    for _ in range(2):
        ops.project_commensurate(v, dims, [0, 1], 2, qb, None, pair_major=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        ops.project_commensurate(v, dims, [0, 1], 2, qb, None, pair_major=True)
    b.record(); torch.cuda.synchronize()
    print("skip=%s threads=%s  %.3f ms" % (os.environ.get("DP_PF_SKIP", "0"), os.environ.get("DP_PF_THREADS", "auto"), a.elapsed_time(b) / 5), flush=True)
else:
    for mask in [int(m) for m in os.environ.get("PF_MASKS", "0,1,2,4,3,5,6,7").split(",")]:
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, DP_PF_SKIP=str(mask)))
