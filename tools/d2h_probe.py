#!/usr/bin/env python
"""How fast can R ranks copy device -> pinned host memory at the same time on this box?  (torchrun, one rank per GPU)

For R in {1, 2, 4, 8} active ranks and three kinds of pinned host memory — cudaHostAlloc (torch pin_memory),
cudaHostRegister of a malloc'ed buffer, cudaHostRegister of a 2 MiB-aligned MADV_HUGEPAGE buffer — time repeated
1 GiB copies, all active ranks between the same two barriers.  Explains the e2e line of bench.py at N > 1."""
import ctypes, json, mmap, os, sys, time
import torch
import torch.distributed as dist

def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    os.environ.setdefault('NCCL_NVLS_ENABLE', '0')
    dist.init_process_group('nccl', device_id=dev)
    nbytes = 1 << 30
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev).fill_(rank + 1)
    cudart = torch.cuda.cudart()
    bufs = {}
    bufs['cudaHostAlloc'] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    plain = torch.empty(nbytes, dtype=torch.uint8)
    plain.fill_(0)
    assert int(cudart.cudaHostRegister(plain.data_ptr(), nbytes, 0)) == 0
    bufs['malloc+cudaHostRegister'] = plain
    libc = ctypes.CDLL('libc.so.6', use_errno=True)
    mm = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    aligned = (base + (2 << 20) - 1) & ~((2 << 20) - 1)
    libc.madvise(ctypes.c_void_p(aligned), ctypes.c_size_t(nbytes), 14)       # MADV_HUGEPAGE
    huge = torch.frombuffer((ctypes.c_char * nbytes).from_address(aligned), dtype=torch.uint8)
    huge.fill_(0)
    assert int(cudart.cudaHostRegister(huge.data_ptr(), nbytes, 0)) == 0
    bufs['hugepage+cudaHostRegister'] = huge
    try:
        thp = open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip()
    except OSError:
        thp = '?'
    rows = []
    for name, dst in bufs.items():
        for active in (1, 2, 4, 8):
            if active > world:
                continue
            dst.copy_(src, non_blocking=True); torch.cuda.synchronize(dev)
            dist.barrier()
            t0 = time.perf_counter()
            if rank < active:
                for _ in range(4):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0 if rank < active else 0.0
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                gbs = 4 * nbytes / float(t.item()) / 1e9
                rows.append({'memory': name, 'ranks_copying': active, 'gbs_per_gpu': gbs, 'gbs_total': gbs * active})
    if rank == 0:
        print(json.dumps({'transparent_hugepage': thp, 'cpus': len(os.sched_getaffinity(0)), 'rows': rows}))
    dist.destroy_process_group()

if __name__ == '__main__':
    main()
