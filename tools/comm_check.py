"""bbbp_comm_* on N ranks (torchrun --nproc-per-node N tools/comm_check.py): the C ABI's two NCCL exchanges -- the score
all-gather of sharded screening and the gradient average of replica training -- against torch.distributed's own collectives.
The ncclUniqueId travels from rank 0 to the others through torch.distributed (any transport would do)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bbbp_b200
from bbbp_b200 import _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.lib
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = (ctypes.c_char * 128)()
    _lib.check(lib.bbbp_comm_unique_id(buf), "comm_unique_id")
    uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
uid = uid.to(dev)
dist.broadcast(uid, 0)
raw = (ctypes.c_char * 128).from_buffer_copy(bytes(uid.cpu().numpy().tobytes()))
comm = ctypes.c_void_p()
_lib.check(lib.bbbp_comm_init_rank(ctypes.byref(comm), world, raw, rank), "comm_init_rank")
stream = torch.cuda.current_stream().cuda_stream

n = 1 << 20                                   # scores per rank
torch.manual_seed(100 + rank)
mine = torch.randn(n, device=dev)
got = torch.empty(world * n, device=dev)
_lib.check(lib.bbbp_comm_gather_scores(comm, mine.data_ptr(), got.data_ptr(), n, stream), "gather_scores")
want = torch.empty(world * n, device=dev)
dist.all_gather_into_tensor(want, mine)
grads = torch.randn(13_464_087, device=dev)   # the flat gradient buffer of MixedInputModel(167, 128)
ref = grads.clone()
_lib.check(lib.bbbp_comm_average_gradients(comm, grads.data_ptr(), grads.numel(), stream), "average_gradients")
dist.all_reduce(ref, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
ok_gather, ok_avg = bool(torch.equal(got, want)), bool(torch.allclose(grads, ref, rtol=0, atol=1e-6))
# device time of the two exchanges
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
t_gather = timed(lambda: lib.bbbp_comm_gather_scores(comm, mine.data_ptr(), got.data_ptr(), n, stream))
t_avg = timed(lambda: lib.bbbp_comm_average_gradients(comm, grads.data_ptr(), grads.numel(), stream))
_lib.check(lib.bbbp_comm_destroy(comm), "comm_destroy")
flags = torch.tensor([int(ok_gather), int(ok_avg)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"ranks": world, "gather_scores_equals_all_gather": bool(flags[0]), "average_gradients_equals_all_reduce_avg": bool(flags[1]),
                      "gather_ms_4MB_per_rank": t_gather, "average_ms_53.9MB": t_avg,
                      "allreduce_bus_GBps": 2 * (world - 1) / world * grads.numel() * 4 / (t_avg * 1e-3) / 1e9}))
dist.destroy_process_group()
sys.exit(0 if bool(flags.min()) else 1)
