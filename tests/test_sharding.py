"""N>1 path on the CPU: two gloo ranks, each steps its share of a 5-column sweep through the CTA emulator, rank 0 gathers
the per-column diagnostics; the result must equal the single-rank run column by column."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from aiolos_b200.sharding import columns_of_rank, gather_diagnostics
from conftest import GOLDEN, ROOT


def test_partition_is_a_permutation():
    for n, w in ((4096, 8), (5, 2), (3, 4)):
        cols = sorted(c for r in range(w) for c in columns_of_rank(n, r, w))
        assert cols == list(range(n))


def test_gather_world1():
    out = gather_diagnostics({"t": np.arange(4.0)}, 4, 0, 1)
    assert out["t"].tolist() == [0, 1, 2, 3]


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r})
    import aiolos_b200 as A
    from aiolos_b200.sharding import columns_of_rank, gather_diagnostics
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_total = 5
    uv = [1e31 * 10 ** (-2 + 4 * k / (n_total - 1)) for k in range(n_total)]
    cols = columns_of_rank(n_total, rank, world)
    ens = A.Ensemble.from_dump({init!r}, ncols=len(cols), lib={lib!r},
                               overrides=[(i, "PARI_UVSTAR", repr(uv[c])) for i, c in enumerate(cols)])
    ens.advance(12)
    st = ens.status_all()
    local = {{"t": st["t"], "steps": st["steps"].astype(float),
              "dS_top": np.array([ens.get("dS", i)[ens.num_cells - 3] for i in range(len(cols))])}}
    out = gather_diagnostics(local, n_total, rank, world)
    if rank == 0:
        json.dump({{k: v.tolist() for k, v in out.items()}}, open({out!r}, "w"))
    dist.destroy_process_group()
""")


def test_two_ranks_match_single_rank(emu_lib, tmp_path):
    import aiolos_b200 as A
    init = os.path.join(GOLDEN, "c4_dynamic3_euv_init.bin")
    out = str(tmp_path / "gathered.json")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, init=init, lib=emu_lib, out=out))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    got = json.load(open(out))
    n_total = 5
    uv = [1e31 * 10 ** (-2 + 4 * k / (n_total - 1)) for k in range(n_total)]
    ens = A.Ensemble.from_dump(init, ncols=n_total, lib=emu_lib, overrides=[(k, "PARI_UVSTAR", repr(uv[k])) for k in range(n_total)])
    ens.advance(12)
    st = ens.status_all()
    assert got["steps"] == [12.0] * n_total
    assert np.array_equal(np.array(got["t"]), st["t"])
    ds = np.array([ens.get("dS", k)[ens.num_cells - 3] for k in range(n_total)])
    assert np.array_equal(np.array(got["dS_top"]), ds)
    assert (np.diff(ds) > 0).all()
