"""world_size-2 gloo test of the multi-GPU plumbing: shard → (oracle WBC as stand-in compute) → one all-gather of the torque
buffer; result equals the single-process result in original robot order."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from qm_control_b200 import parallel, synthetic
    from _oracle import Oracle
    r, w, _ = parallel.init_distributed(backend="gloo")
    lo, hi = parallel.shard_range(total, r, w)
    prob, wbc = synthetic.make_batch(np.arange(lo, hi), config=5)
    o = Oracle(); x, u, mode = synthetic.nominal_wbc_inputs(prob, o.model_info()["mass"])
    cmd, _ = o.wbc_update_batch(x, u, wbc["rbd"], mode, wbc["period"], wbc["time"], u, nthreads=2)
    tau = torch.from_numpy(cmd[:, 36:].copy())
    full = parallel.allgather_torque(tau, total, r, w)
    t = parallel.max_over_ranks(float(r + 1), "cpu")
    if r == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy()); np.save(os.path.join(out_dir, "tmax.npy"), np.array([t]))
    torch.distributed.destroy_process_group()


def test_allgather_torque_two_ranks(tmp_path, oracle):
    from qm_control_b200 import synthetic
    total = 7   # uneven shards (4 + 3)
    mp.spawn(_worker, args=(2, 29533, total, str(tmp_path)), nprocs=2, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    prob, wbc = synthetic.make_batch(np.arange(total), config=5)
    x, u, mode = synthetic.nominal_wbc_inputs(prob, oracle.model_info()["mass"])
    cmd, _ = oracle.wbc_update_batch(x, u, wbc["rbd"], mode, wbc["period"], wbc["time"], u, nthreads=2)
    np.testing.assert_array_equal(gathered, cmd[:, 36:])
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0
