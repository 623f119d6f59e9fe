"""world_size-2 gloo test (CPU) of the k-point sharding used on N>1 GPUs: every k-point solved
exactly once, contiguous chunks, results back in path order.  The solve itself is replaced by
the exact empty-lattice spectrum so no GPU is needed."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mfem_bravais_b200 as m
    from oracle.bloch_oracle import Lattice, empty_lattice_eigs
    lat = m.BravaisLattice("FCC")
    ks = m.k_path(lat, ["Gamma", "X", "W", "L", "Gamma"], 8)[:13]      # ragged: 13 points on 2 ranks
    olat = Lattice("FCC")
    solved = []

    def solve(k):
        solved.append(tuple(np.round(k, 12)))
        return empty_lattice_eigs(olat, k, 6)

    out = m.sharded_sweep(solve, ks, 6, dist)
    np.save(os.path.join(outdir, "out%d.npy" % rank), out)
    np.save(os.path.join(outdir, "n%d.npy" % rank), np.array([len(solved)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    sys.path.insert(0, ROOT)
    import mfem_bravais_b200 as m
    for n in (0, 1, 7, 32, 256):
        for w in (1, 2, 3, 4, 8):
            chunks = [m.shard_kpoints(n, w, r) for r in range(w)]
            assert chunks[0][0] == 0 and chunks[-1][1] == n
            assert all(chunks[i][1] == chunks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in chunks]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_sweep(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "out0.npy"), np.load(tmp_path / "out1.npy")
    assert np.array_equal(a, b) and a.shape == (13, 6)
    assert int(np.load(tmp_path / "n0.npy")[0]) == 7 and int(np.load(tmp_path / "n1.npy")[0]) == 6
    sys.path.insert(0, ROOT)
    import mfem_bravais_b200 as m
    from oracle.bloch_oracle import Lattice, empty_lattice_eigs
    ks = m.k_path(m.BravaisLattice("FCC"), ["Gamma", "X", "W", "L", "Gamma"], 8)[:13]
    ref = np.array([empty_lattice_eigs(Lattice("FCC"), k, 6) for k in ks])
    assert np.allclose(a, ref)
