"""Probe: does CUDA IPC (cudaIpcGetMemHandle / cudaIpcOpenMemHandle) work between the ranks of one box?
Run under torch.distributed.run with >= 2 ranks.  Rank r writes a pattern into rank (r+1)'s buffer through the
imported pointer; every rank checks what its left neighbour wrote."""
import ctypes as C
import os
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo", rank=rank, world_size=world)
rt = C.CDLL("libcudart.so.12")
ptr = C.c_void_p()
assert rt.cudaSetDevice(local) == 0
assert rt.cudaMalloc(C.byref(ptr), C.c_size_t(1 << 20)) == 0
assert rt.cudaMemset(ptr, 0, C.c_size_t(1 << 20)) == 0
handle = C.create_string_buffer(64)
rc = rt.cudaIpcGetMemHandle(handle, ptr)
print(f"rank {rank}: cudaIpcGetMemHandle rc={rc}", flush=True)
handles = [None] * world
dist.all_gather_object(handles, handle.raw)
right = (rank + 1) % world
peer = C.c_void_p()
class H(C.Structure):
    _fields_ = [("b", C.c_char * 64)]
h = H(); C.memmove(C.byref(h), handles[right], 64)
rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), H, C.c_uint]
rc = rt.cudaIpcOpenMemHandle(C.byref(peer), h, C.c_uint(1))
print(f"rank {rank}: cudaIpcOpenMemHandle(right={right}) rc={rc} peer={peer.value}", flush=True)
can = C.c_int(0)
rt.cudaDeviceCanAccessPeer(C.byref(can), local, right % torch.cuda.device_count())
print(f"rank {rank}: canAccessPeer={can.value}", flush=True)
if rc == 0:
    src = torch.full((1024,), rank + 100, dtype=torch.int32, device="cuda")
    rc2 = rt.cudaMemcpy(peer, C.c_void_p(src.data_ptr()), C.c_size_t(4096), C.c_int(3))
    torch.cuda.synchronize()
    dist.barrier()
    out = torch.zeros(1024, dtype=torch.int32, device="cuda")
    rt.cudaMemcpy(C.c_void_p(out.data_ptr()), ptr, C.c_size_t(4096), C.c_int(3))
    torch.cuda.synchronize()
    left = (rank - 1) % world
    ok = bool((out == left + 100).all().item())
    print(f"rank {rank}: memcpy rc={rc2}, received from left {left}: {'OK' if ok else 'MISMATCH'} ({out[:2].tolist()})", flush=True)
dist.barrier()
