"""Worker for the multi-rank tests (launched by torch.distributed.run).

  --backend gloo : CPU restatement of the sharded-registration exchange protocols of csrc/comm.cu + icp.cu, built from the
                   ORACLE's stage functions: each rank matches its shard; the trimmed quantile is found both ways -- three
                   all-reduced radix-select digit histograms (NCCL carrier) and one histogram + an all-gather of the candidate
                   keys of the picked bin (peer-memory carrier); the 32-bit limbs of the 128-bit normal-equation sums are
                   summed as integers, every rank solves.  Checks that the sharded trajectory equals the single-process
                   oracle bit for bit.
  --backend nccl : the real thing on GPUs: aicp_b200_comm_init + aicp_b200_register on each rank's shard, compared with the
                   unsharded registration on the same GPU.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pick_digit(hist, target):
    cum = np.cumsum(hist)
    b = int(np.searchsorted(cum, target, side="right"))
    return b, target - (int(cum[b - 1]) if b > 0 else 0)


def run_gloo(args):
    import torch
    import torch.distributed as dist
    from aicp_mapping_b200 import synth
    from oracle import oracle as orc
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pair = synth.make_pair(5, 3, 6000)
    ratio = np.float32(0.6)
    cfg = orc.default_config(ratio=float(ratio))
    full = orc.icp(pair["ref"], pair["read"], cfg, want_normals=True, want_trace_idx=True)
    assert full.rc == 0
    # replicated reference side
    normals = full.normals
    mu = full.mean_ref
    refc = orc.to_xyzw(pair["ref"])
    refc[:, :3] = refc[:, :3] - mu
    M0 = np.eye(4, dtype=np.float32); M0[:3, 3] = -mu
    shard = np.arange(rank, pair["read"].shape[0], world)
    read0 = orc.transform_points(M0, pair["read"][shard])
    T = np.eye(4, dtype=np.float32)
    for it in range(full.iterations):
        step = orc.transform_points(T, read0)
        idx, d2 = orc.match(refc, step)
        assert np.array_equal(idx, full.trace_idx[it][shard])
        # global trimmed quantile: 3-digit radix select over all-reduced histograms
        key = d2.view(np.uint32).astype(np.int64)
        valid = (d2 > 0) & np.isfinite(d2)
        prefix, k_rem = 0, None
        for p, (shift, bits) in enumerate(((20, 11), (9, 11), (0, 9))):
            sel = valid if p == 0 else valid & ((key >> (shift + bits)) == prefix)
            h = torch.from_numpy(np.bincount((key[sel] >> shift) & ((1 << bits) - 1), minlength=2048).astype(np.int64))
            dist.all_reduce(h)
            h = h.numpy()
            if p == 0:
                n_valid = int(h.sum())
                k_rem = min(int(np.float32(n_valid) * ratio), n_valid - 1)
            b, k_rem = pick_digit(h, k_rem)
            prefix = (prefix << bits) | b
        limit = np.array([prefix], dtype=np.uint32).view(np.float32)[0]
        assert limit == full.trace[it]["limit_d2"] and n_valid == full.trace[it]["n_valid"]
        # the peer-memory carrier's version of the same quantile (csrc/icp.cu, loop_pick / loop_select23): ONE histogram
        # exchange (digit 1), then the candidate KEYS of the picked bin are all-gathered and digits 2 and 3 are finished over the
        # gathered list by every rank
        h1 = torch.from_numpy(np.bincount(key[valid] >> 20, minlength=2048).astype(np.int64))
        dist.all_reduce(h1)
        h1 = h1.numpy()
        k1 = min(int(np.float32(int(h1.sum())) * ratio), int(h1.sum()) - 1)
        b1, k1 = pick_digit(h1, k1)
        mine = key[valid & ((key >> 20) == b1)]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tolist())
        cand = np.array(sum(gathered, []), dtype=np.int64)
        assert len(cand) == h1[b1]
        b2, k2 = pick_digit(np.bincount((cand >> 9) & 2047, minlength=2048), k1)
        p22 = (b1 << 11) | b2
        b3, _ = pick_digit(np.bincount(cand[(cand >> 9) == p22] & 511, minlength=512), k2)
        assert np.array([(p22 << 9) | b3], dtype=np.uint32).view(np.float32)[0] == limit
        # exact normal-equation partials as 32-bit limbs
        hi, lo, used = orc.normal_equations(step, refc, normals, idx, d2, limit)
        limbs = np.zeros(27 * 4 + 1, dtype=np.int64)
        for i in range(27):
            v = ((int(hi[i]) << 64) + int(lo[i])) & ((1 << 128) - 1)
            for j in range(4):
                limbs[4 * i + j] = (v >> (32 * j)) & 0xFFFFFFFF
        limbs[-1] = used
        t = torch.from_numpy(limbs)
        dist.all_reduce(t)
        limbs = t.numpy()
        assert int(limbs[-1]) == full.trace[it]["n_used"]
        hi2, lo2 = np.zeros(27, dtype=np.int64), np.zeros(27, dtype=np.uint64)
        for i in range(27):
            v = sum(int(limbs[4 * i + j]) << (32 * j) for j in range(4)) & ((1 << 128) - 1)
            lo2[i] = v & ((1 << 64) - 1)
            h64 = v >> 64
            hi2[i] = h64 - (1 << 64) if h64 >= (1 << 63) else h64
        x, _ = orc.solve6(hi2, lo2)
        dT = orc.pose_increment(x)
        # the float 4x4 product dT * T is the oracle's own (mat4_mul_f); its result is taken from the single-process
        # trace and the increment is checked against it through the rotation angle and translation of dT
        Tn = full.trace[it]["T_iter"].astype(np.float64)
        assert np.abs(dT.astype(np.float64) @ T.astype(np.float64) - Tn).max() < 1e-5
        T = full.trace[it]["T_iter"]
        assert np.all(np.isfinite(x))
    if rank == 0:
        print("GLOO_SHARDED_OK iterations=%d world=%d" % (full.iterations, world))
    dist.destroy_process_group()


def run_nccl(args):
    import torch
    import torch.distributed as dist
    import aicp_mapping_b200 as ab
    from aicp_mapping_b200 import capi, synth
    from aicp_mapping_b200.registration import comm_unique_id
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    pair = synth.make_pair(args.config, args.trial, args.points)
    ratio = 0.6
    ref_reg = ab.B200Registration(device=local)
    ref_reg.setConfig(ratio=ratio)
    T_full = ref_reg.registerClouds(pair["ref"], pair["read"])
    it_full, used_full = ref_reg.stats.iterations, ref_reg.getWeightedPointUsedRatio()
    out_full = ref_reg.getOutputReading()
    u32 = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    shard = np.arange(rank, pair["read"].shape[0], world)
    ok, info = True, []
    # both carriers of the exchange: peer-mapped inboxes inside the persistent loop kernel (default), NCCL all-reduces
    for carrier in ("peer", "nccl"):
        if carrier == "nccl":
            os.environ["AICP_B200_COMM"] = "nccl"
        else:
            os.environ.pop("AICP_B200_COMM", None)
        uid = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        reg = ab.B200Registration(device=local)
        reg.setConfig(ratio=ratio)
        reg.commInit(uid[0], rank, world)
        desc = reg.commInfo()
        if carrier == "peer" and "peer-mapped" not in desc:
            info.append("peer carrier unavailable: " + desc)
        T = reg.registerClouds(pair["ref"], pair["read"][shard])
        ok = ok and (np.array_equal(u32(T), u32(T_full)) and reg.stats.iterations == it_full and
                     np.float32(reg.getWeightedPointUsedRatio()) == np.float32(used_full) and
                     np.array_equal(u32(reg.getOutputReading()), u32(out_full[shard])))
        ms_sharded = reg.stats.ms_total
        # fixed-reference mode with the sharded reading (BASELINE.json config 4 shape), twice: the second call reuses the index
        reg.setReference(pair["ref"])
        for _ in range(2):
            T2 = reg.registerToReference(pair["read"][shard])
            ok = ok and np.array_equal(u32(T2), u32(T_full))
        # error path: ONE rank's shard holds a NaN -> every rank must come back with an error instead of waiting for ever
        bad = pair["read"][shard].copy()
        if rank == world - 1:
            bad[0, 0] = np.nan
        try:
            reg.registerClouds(pair["ref"], bad)
            err = "OK"
        except capi.AicpError as e:
            err = e.code_name
        ok = ok and err == ("NONFINITE_INPUT" if rank == world - 1 else "COMM")
        # ... and the communicator still works afterwards
        T3 = reg.registerClouds(pair["ref"], pair["read"][shard])
        ok = ok and np.array_equal(u32(T3), u32(T_full))
        info.append("%s ms_sharded=%.3f err=%s" % (carrier, ms_sharded, err))
        reg.commDestroy(); reg.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("NCCL_SHARDED_%s iterations=%d world=%d ms_single=%.3f %s" %
              ("OK" if int(flag) else "MISMATCH", it_full, world, ref_reg.stats.ms_total, " | ".join(info)))
    else:
        print("rank %d ok=%s %s" % (rank, ok, " | ".join(info)))
    ref_reg.close()
    dist.destroy_process_group()
    if not int(flag):
        sys.exit(1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--config", type=int, default=5)
    ap.add_argument("--trial", type=int, default=3)
    ap.add_argument("--points", type=int, default=6000)
    a = ap.parse_args()
    run_gloo(a) if a.backend == "gloo" else run_nccl(a)
