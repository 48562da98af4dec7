"""CPU, world_size 2 over gloo: the multi-GPU MSM / Groth16 exchange.  Each rank produces the partials of
its point-range shard (here with the oracle standing in for the GPU), the ranks all_gather the raw partial bytes, and
the product's host-side combine (sb_host_sum_partials / sb_host_groth16_finish) must reproduce the unsharded result."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BN = O.BN254


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _partial_of(lib, grp, jac_bytes):
    """oracle Jacobian result -> product partial (XYZZ) via affine."""
    aff = np.frombuffer(O.g_to_affine(BN, grp, jac_bytes), np.uint8).copy()
    out = np.empty(lib.sb_host_partial_bytes(0, grp), np.uint8)
    assert lib.sb_host_partial_from_affine(0, grp, _ptr(aff), _ptr(out)) == 0
    return out


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snarkjs_b200 import _native as N
    lib = N.lib()
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "groth16_case.npz")))
    zkey, wt = g["zkey"].tobytes(), g["wtns"].tobytes()
    zdata, zsecs = O.read_binfile(zkey, "zkey", 2)
    zk = O.read_zkey_header(zdata, zsecs)
    _, W = O.read_wtns(wt)
    Wb = np.frombuffer(W, np.uint8)
    ci = O.CURVES[BN]
    n, nv, npub = zk["domainSize"], zk["nVars"], zk["nPublic"]
    # H scalars (replicated NTT chain) from the oracle
    A_T, B_T, C_T = O.build_abc(BN, bytes(O.section(zdata, zsecs, 4)), W, n)
    inc = O.fr_root(BN, zk["power"] + 1)
    odd = [O.fr_fft(BN, O.fr_batch_apply_key(BN, O.fr_fft(BN, X, inverse=True), ci.fr_to_mont(1), inc)) for X in (A_T, B_T, C_T)]
    P = O.qap_join_abc(BN, *odd)
    # C bases padded like the product does, so every witness MSM shares the index range
    secC = bytes(64 * (npub + 1)) + bytes(O.section(zdata, zsecs, 8))

    def rng(total):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        lib.sb_shard_range(total, rank, world, ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value
    wlo, wcnt = rng(nv)
    hlo, hcnt = rng(n)
    parts = []
    for grp, sec, lo, cnt, sc in ((1, bytes(O.section(zdata, zsecs, 5)), wlo, wcnt, Wb), (1, bytes(O.section(zdata, zsecs, 6)), wlo, wcnt, Wb),
                                  (1, secC, wlo, wcnt, Wb), (1, bytes(O.section(zdata, zsecs, 9)), hlo, hcnt, P),
                                  (2, bytes(O.section(zdata, zsecs, 7)), wlo, wcnt, Wb)):
        sz = 64 * grp
        jac = O.multiexp_affine(BN, grp, sec[lo * sz:(lo + cnt) * sz], sc[lo * 32:(lo + cnt) * 32], 2)
        parts.append(_partial_of(lib, grp, jac))
    mine = torch.from_numpy(np.concatenate(parts))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)                      # the one exchange step of the path
    allp = torch.cat(gathered).numpy()
    r, s = ci.fr_to_mont(123456789), ci.fr_to_mont(987654321)
    proof = np.empty(256, np.uint8)
    rc = lib.sb_host_groth16_finish(0, zk["vk_alpha_1"], zk["vk_beta_1"], zk["vk_beta_2"], zk["vk_delta_1"], zk["vk_delta_2"],
                                    _ptr(allp), world, r, s, _ptr(proof))
    assert rc == 0
    # plain MSM partial sum too
    pb = lib.sb_host_partial_bytes(0, 1)
    a_parts = np.concatenate([gathered[i].numpy()[:pb] for i in range(world)])
    jac = np.empty(96, np.uint8)
    assert lib.sb_host_sum_partials(0, 1, _ptr(a_parts), world, _ptr(jac)) == 0
    q.put((rank, proof.tobytes(), jac.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_groth16_exchange_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "groth16_case.npz")))
    ci = O.CURVES[BN]
    r, s = ci.fr_to_mont(123456789), ci.fr_to_mont(987654321)
    oproof, _pub, parts = O.groth16_prove(g["zkey"].tobytes(), g["wtns"].tobytes(), r, s, return_parts=True)
    for rank, proof, jac in res:
        aff = (proof[:64], proof[64:192], proof[192:256])
        assert O.proof_to_object(ci, aff) == oproof, rank
        assert O.g_to_affine(BN, 1, jac) == parts["msm_affine"]["A"], rank
