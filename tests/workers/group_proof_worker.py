"""torchrun worker: ONE proof computed jointly by WORLD_SIZE GPUs (zk_group) must have the bytes of the CPU prover's proof.
Run as: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/workers/group_proof_worker.py <shape> <k> [expect_p2p]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import zkdcap_b200 as zk
from oracle import oracle as orc
from oracle import prover as op
from tests import prover_common as pc


def main():
    shape, k = sys.argv[1], int(sys.argv[2])
    expect_p2p = len(sys.argv) > 3 and sys.argv[3] == "1"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = zk.Context(local)
    case = pc.make_case(shape, k, 7)
    p = pc.params_for(k)
    params = zk.ParamsKZG(k, p["g"], p["g_lagrange"], ctx, precompute=True)
    key = zk.ProverKey(params, case["circ"], case["fixed"], case["mapping"])
    plain = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    uid = [zk.Group.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    grp = zk.Group(ctx, rank, world, uid[0])
    p2p = bool(ctx.lib.zk_group_p2p_enabled(grp.h))
    key.set_group(grp)
    proofs = [key.create_proof(case["advice"], case["instances_limbs"], case["randoms"]) for _ in range(3)]
    key.set_group(None)
    again = key.create_proof(case["advice"], case["instances_limbs"], case["randoms"])
    ok = all(pr == plain for pr in proofs) and again == plain and (p2p or not expect_p2p)
    if rank == 0:
        opk = op.keygen(orc, p, case["circ"], case["fixed"], case["mapping"], case["circ"].transcript_repr)
        want = op.create_proof(orc, p, opk, case["circ"], case["advice"], case["instances_int"], case["randoms"])
        ok = ok and plain == want
    # ONE MSM sharded by base range (zk_msm_sharded_dev): every rank holds its slice of the bases as its resident SRS
    import numpy as np
    km = 12
    n = 1 << km
    per = n // world
    bases, scal = orc.gen_points(0x5A0 + km, n), orc.gen_scalars(0x5A1 + km, n)
    ks = km - (world.bit_length() - 1)
    sl = slice(rank * per, (rank + 1) * per)
    zk.ParamsKZG(ks, bases[sl], bases[sl], ctx, precompute=True)
    d_s = torch.from_numpy(np.ascontiguousarray(scal[sl]).view(np.int64)).cuda()
    d_o = torch.zeros((1, 12), dtype=torch.int64, device="cuda")
    for _ in range(3):                                      # repeated: the exchange windows alternate
        grp.msm_sharded_dev(d_s.data_ptr(), per, d_o.data_ptr())
    torch.cuda.synchronize()
    got = d_o.cpu().numpy().view(np.uint64).reshape(12)
    msm_ok = bool(np.array_equal(orc.g1_to_affine(got), orc.g1_to_affine(orc.best_multiexp(scal, bases))))
    ok = ok and msm_ok
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    grp.close()
    if rank == 0:
        print("GROUP_PROOF world=%d p2p=%s ok=%s" % (world, p2p, bool(t.item())))
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
