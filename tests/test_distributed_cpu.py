"""N>1 host logic on CPU (gloo, world_size 2): shard ranges, the all-gather of partial-sum
blobs, and the host-side addition of per-shard partials (bb_point_add runs on the host).  The
per-shard sums themselves come from the CPU oracle here; on the GPU box the same plumbing carries
the device results (bench.py --gpus N, tests/test_gpu_parity.py's sharded finalize)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.environ["BB_ROOT"])
    import bellman_b200 as bb
    from bellman_b200.distributed import shard_range, all_gather_partials, max_over_ranks
    from oracle import o1

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 301
    bases = o1.g1_fixed_mul(o1.fr_random(1, n))
    ex = o1.fr_random(2, n)
    lo, hi = shard_range(n, rank, world)
    rc, part = o1.multiexp(1, bases[lo:hi], 0, None, ex[lo:hi])      # this rank's base range
    assert rc == 0
    blob = bytearray(bb.PARTIALS_BYTES)
    blob[0:96] = part.tobytes()                                       # slot 0 = the "h" partial
    blob[96] = rank                                                   # marker to check ordering
    got = all_gather_partials(bytes(blob))
    assert [g[96] for g in got] == list(range(world))
    total = np.zeros((1, 12), np.uint64)
    for g in got:
        total = bb.point_add(bb.G1, total, np.frombuffer(g[:96], dtype=np.uint64).reshape(1, 12))
    rc, want = o1.multiexp(1, bases, 0, None, ex)
    assert rc == 0 and np.array_equal(total, want)
    assert max_over_ranks(float(rank)) == world - 1
    ranges = [shard_range(n, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    # error on one rank: every rank raises, nobody hangs in the collective
    import bellman_b200.distributed as D
    def fake_partials(assignment, params, device_ptrs=None):
        if rank == 1:
            raise bb.UnexpectedIdentity("[bb_status 2] identity in the CRS shard of rank 1")
        return bytes(bb.PARTIALS_BYTES)
    bb.prove_partials = fake_partials
    try:
        D.create_proof_sharded(None, None, None, 1, 2, split_h=False)
        raise SystemExit("expected UnexpectedIdentity")
    except bb.UnexpectedIdentity as e:
        assert "rank 1" in str(e) and "identity in the CRS shard of rank 1" in str(e)   # the failing rank's own text, on every rank
    dist.destroy_process_group()
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"ok.{rank}"), "w").write("ok")   # one file per rank: stdout of the two ranks can interleave
''')


def test_sharded_partials_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, BB_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert (tmp_path / "ok.0").exists() and (tmp_path / "ok.1").exists(), res.stdout + res.stderr


SHARDED_PROVE = textwrap.dedent('''
    import os, sys, random
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.environ["BB_ROOT"]); sys.path.insert(0, os.path.join(os.environ["BB_ROOT"], "tests"))
    import bellman_b200 as bb
    bb.LIB_PATH, bb._lib = os.environ["BB_EMU_LIB"], None            # host-fiber execution of the product sources
    from bellman_b200.distributed import create_proof_sharded
    from oracle import o1

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    R = o1.FR_MODULUS
    rng = random.Random(5)
    mc = o1.Mimc(40, seed=9)
    mc.set_toxic([rng.randrange(1, R) for _ in range(5)])
    mc.generate()
    w = mc.witness()
    asg = bb.ProvingAssignment(w["a"], w["b"], w["c"], w["inputs"], w["aux"], w["a_aux_density"], w["b_input_density"], w["b_aux_density"])
    worker = bb.Worker(0)
    mine = bb.Parameters(worker, mc.export_params(), shard_index=rank, shard_count=world)
    r, s = rng.randrange(R), rng.randrange(R)
    for split_h in (True, False):                 # H pipeline divided by polynomial across the ranks / replicated
        proof = create_proof_sharded(asg, mine, mine, r, s, split_h=split_h)
        if rank == 0:
            assert proof == mc.prove(r, s) == mc.expected_proof(r, s), split_h
        else:
            assert proof is None
    if world == 2:                                 # (once is enough: the three-rank run covers the 3-way H split)
        # the MSM form is chosen collectively: every form proves, every form is eligible here, all ranks agree
        from bellman_b200.distributed import autotune_sharded
        rep = autotune_sharded(asg, mine, r, s, reps=1)
        names = bb.tuning_names()
        assert len(rep["ms"]) == len(names) and all(t > 0 for t in rep["ms"]) and rep["ms"][rep["chosen"]] == min(rep["ms"]), rep
        picks = [None] * world
        dist.all_gather_object(picks, rep["chosen"])
        assert len(set(picks)) == 1
        for index in [rep["chosen"]] + list(range(len(names))):
            if index != rep["chosen"]:
                mine.apply_tuning(index)
            proof = create_proof_sharded(asg, mine, mine, r, s)
            assert proof is None if rank else proof == mc.expected_proof(r, s), names[index]
        mine.apply_tuning(0)
        # bench.py's gate in front of the collective tuner: the per-rank tuner child (stubbed here: there is no CUDA device to
        # give a child) must survive on EVERY rank, otherwise all ranks stay on the default form
        import importlib.util, types
        spec = importlib.util.spec_from_file_location("bench", os.path.join(os.environ["BB_ROOT"], "bench.py"))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
        args = types.SimpleNamespace(autotune=1, autotune_reps=1)
        calls = []
        def child(a, log_n, shard=(0, 1), reps=None):
            calls.append(shard)
            return {"chosen": 1, "ms": [1.0], "error": None}
        bench.tune_in_child = child
        rep = bench.tune_sharded_gated(args, 7, rank, world, asg, mine, r, s, None)
        assert calls == [(rank, world)] and rep["error"] is None and all(t > 0 for t in rep["ms"]) and rep["ms"][rep["chosen"]] == min(rep["ms"])
        bench.tune_in_child = lambda a, log_n, shard=(0, 1), reps=None: {"chosen": 0, "ms": None, "error": "boom" if rank == world - 1 else None}
        rep = bench.tune_sharded_gated(args, 7, rank, world, asg, mine, r, s, None)
        assert rep["chosen"] == 0 and rep["error"] and ("boom" in rep["error"] or "another rank" in rep["error"])
        mine.apply_tuning(0)
    # a shard with an identity base: every rank raises the same SynthesisError
    bad = mc.export_params()
    if rank == 1:
        bad["l"] = np.array(bad["l"], copy=True); bad["l"].reshape(-1, 12)[-1] = 0
    broken = bb.Parameters(worker, bad, shard_index=rank, shard_count=world)
    try:
        create_proof_sharded(asg, broken, broken, r, s)
        raise SystemExit("expected UnexpectedIdentity")
    except bb.UnexpectedIdentity:
        pass
    worker.close()
    dist.destroy_process_group()
    open(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"ok.{rank}"), "w").write("ok")
''')


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_prove_gloo_emulated_device(tmp_path, emu_lib, world):
    """The whole N>1 path on CPU: gloo ranks, each running the product sources on host fibers
    (tests/native) over its (base range x window) shard; the H pipeline split by polynomial (two ranks:
    a, c | b; three: a | b | c) with broadcasts of the coset evaluations, and replicated; one all-gather,
    finalize on rank 0 with the static terms computed on a host thread meanwhile."""
    script = tmp_path / "worker.py"
    script.write_text(SHARDED_PROVE)
    env = dict(os.environ, BB_ROOT=ROOT, BB_EMU_LIB=emu_lib)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29519 + world), str(script)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout + res.stderr
    assert all((tmp_path / f"ok.{r}").exists() for r in range(world)), res.stdout + res.stderr
