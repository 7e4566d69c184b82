"""N > 1 host logic on CPU: world_size-2 `gloo` processes.

Decoding shards by utterance with no collective; EM all-reduces ONE packed float64 accumulator
(S0 | S1 | S2c | xi_sum | pi_occ | loglik | n_utt -- the layout of lnasr_em_acc_doubles, include/lnasr_b200.h) and
then every rank finalises the same numbers.  The per-rank statistics come from the oracle here (the CUDA kernels
need a GPU; their single-rank parity is tests/test_em_gpu.py) -- what is under test is `lnasr.dist`: the partition,
the reduce hook `train_batch(reduce_fn=...)` receives, and that replicas end bit-identical.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pack(acc):
    return np.concatenate([acc["S0"].ravel(), acc["S1"].ravel(), acc["S2c"].ravel(), acc["xi_sum"].ravel(),
                           acc["pi_occ"].ravel(), [acc["loglik"]], [acc["n_utt"]]])


def _unpack(v, n, m, d):
    o = 0
    out = {}
    for k, shp in (("S0", (n, m)), ("S1", (n, m, d)), ("S2c", (n, m, d)), ("xi_sum", (n, n)), ("pi_occ", (n,))):
        sz = int(np.prod(shp))
        out[k] = v[o:o + sz].reshape(shp)
        o += sz
    out["loglik"], out["n_utt"] = float(v[o]), float(v[o + 1])
    return out


def _make_case():
    from oracle import lnasr_oracle as O
    rng = np.random.default_rng(7)
    n, m, d = 3, 8, 5
    mdl = O.random_gmmhmm(n, m, d, rng)
    utts = [0.6 * rng.standard_normal((int(T), d)) for T in rng.integers(3, 40, size=11)]
    return (n, m, d), mdl, utts


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, "ln-asr_b200"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from lnasr import _capi
    from lnasr.dist import make_reduce_fn, shard_utterances, world as world_fn
    from oracle import lnasr_oracle as O

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert world_fn() == (rank, world)
        (n, m, d), mdl, utts = _make_case()
        mine = shard_utterances([len(u) for u in utts], world, rank)
        acc = np.zeros(_capi.lib().lnasr_em_acc_doubles(n, m, d))          # host helper of the C ABI, no device work
        for i in mine:
            acc += _pack(O.em_accumulate_diag(utts[i], mdl["pi"], mdl["A"], mdl["w"], mdl["mu"], mdl["var"]))
        t = torch.from_numpy(acc)
        make_reduce_fn()(t)                                                # the hook train_batch() is given
        new = O.em_finalize_diag(_unpack(t.numpy(), n, m, d))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), owned=np.asarray(mine), acc=t.numpy(), **new)
    finally:
        dist.destroy_process_group()


def test_em_allreduce_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    from oracle import lnasr_oracle as O
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    (n, m, d), mdl, utts = _make_case()
    # the partition covers every utterance exactly once
    assert sorted(np.concatenate([x["owned"] for x in r]).tolist()) == list(range(len(utts)))
    assert all(len(x["owned"]) > 0 for x in r)
    # replicas are bit-identical after the all-reduce + finalise
    for k in ("acc", "A", "pi", "w", "mu", "var"):
        assert np.array_equal(r[0][k], r[1][k]), k
    # and equal the single-process multi-utterance EM step
    ref, tot = O.em_step_multi_diag(utts, mdl)
    np.testing.assert_allclose(r[0]["acc"], _pack(tot), rtol=1e-12, atol=1e-300)
    for k in ("mu", "var"):
        np.testing.assert_allclose(r[0][k], ref[k], rtol=1e-10)
    with np.errstate(all="ignore"):
        for k in ("A", "pi", "w"):
            np.testing.assert_allclose(np.exp(r[0][k]), np.exp(ref[k]), rtol=1e-10, atol=1e-300)


def test_reduce_fn_is_identity_without_a_process_group():
    import torch
    from lnasr.dist import make_reduce_fn, world
    assert world() == (0, 1)
    t = torch.arange(5, dtype=torch.float64)
    assert torch.equal(make_reduce_fn()(t.clone()), t)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_sharding_is_a_partition_for_any_world(world):
    from lnasr.dist import shard_utterances
    lens = np.random.default_rng(world).integers(1, 500, size=257)
    parts = [shard_utterances(lens, world, r) for r in range(world)]
    assert sorted(i for p in parts for i in p) == list(range(257))
    loads = [int(lens[p].sum()) for p in parts]
    assert max(loads) - min(loads) <= int(lens.max())
