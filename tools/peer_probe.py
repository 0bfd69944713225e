"""Probe (2+ GPUs, torchrun): which peer-memory mechanisms work on this box?
  (a) NCCL transport lines (NCCL_DEBUG=INFO), (b) torch symmetric memory rendezvous + a peer read,
  (c) legacy CUDA IPC handle of a torch allocation opened in the peer process (via torch's own reductions)."""
import os, sys, time, traceback
import torch
import torch.distributed as dist

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
x = torch.ones(1024, device='cuda') * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize()
if rank == 0: print('allreduce ok', float(x[0]), flush=True)
print(f'rank {rank}: can_access_peer', [torch.cuda.can_device_access_peer(lr, j) for j in range(world) if j != lr], flush=True)
# (b) symmetric memory
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(4096, dtype=torch.float32, device=torch.device('cuda', lr))
    t.fill_(rank + 1)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    torch.cuda.synchronize(); dist.barrier()
    print(f'rank {rank}: symm_mem ok, buffer_ptrs={[hex(p) for p in hdl.buffer_ptrs]} signal_pads={[hex(p) for p in hdl.signal_pad_ptrs]} '
          f'pad_size={getattr(hdl, "signal_pad_size", None)} multicast={hex(getattr(hdl, "multicast_ptr", 0) or 0)}', flush=True)
    peer = hdl.get_buffer((rank + 1) % world, (4096,), torch.float32)
    v = float(peer[:8].sum().item())
    print(f'rank {rank}: peer read sum={v} (expect {8 * ((rank + 1) % world + 1)})', flush=True)
    # timing of the handle's own device barrier
    hdl.barrier(channel=0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): hdl.barrier(channel=0)
    e1.record(); torch.cuda.synchronize()
    print(f'rank {rank}: symm barrier {e0.elapsed_time(e1) * 10:.1f} us each', flush=True)
except Exception:
    print(f'rank {rank}: symm_mem FAILED\n' + traceback.format_exc()[-1500:], flush=True)
# NCCL small-message latencies for reference
for n in (77000, 3 * 1024 * 1024 // 4):
    y = torch.ones(n, device='cuda')
    for _ in range(5): dist.all_reduce(y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): dist.all_reduce(y)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f'nccl all_reduce {n * 4} B: {e0.elapsed_time(e1) * 20:.1f} us', flush=True)
g = torch.empty(world * 95250, dtype=torch.int32, device='cuda'); l = torch.ones(95250, dtype=torch.int32, device='cuda')
for _ in range(5): dist.all_gather_into_tensor(g, l)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): dist.all_gather_into_tensor(g, l)
e1.record(); torch.cuda.synchronize()
if rank == 0: print(f'nccl all_gather {95250 * 4} B per rank: {e0.elapsed_time(e1) * 20:.1f} us', flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
