import sys, time, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nn_distributed_training_b200.ops import load_ext
ext = load_ext(required=True)
print('affinity', len(os.sched_getaffinity(0)))
L,B,P,M = 10,64,2,22000
x = torch.randint(0,255,(L*M,784),dtype=torch.uint8); y = torch.randint(0,10,(L*M,),dtype=torch.int64)
for threads in (1,4,8):
    nslots=64
    xp = torch.empty(nslots,P,L,B,784,dtype=torch.uint8); yp=torch.empty(nslots,P,L,B,dtype=torch.int64); bp=torch.empty(nslots,P,L,dtype=torch.int32)
    xp.zero_(); yp.zero_()
    t=time.perf_counter()
    ld = ext.HostBatchLoader(x.data_ptr(), y.data_ptr(), 784, [i*M for i in range(L)], [M]*L, [0]*L, B, P, 0, 0,
        [xp[s].data_ptr() for s in range(nslots)],[yp[s].data_ptr() for s in range(nslots)],[bp[s].data_ptr() for s in range(nslots)], threads)
    while ld.rounds_assembled() < nslots: time.sleep(0.0005)
    # rounds_assembled counts claimed rounds; wait for all ready by acquiring all
    for _ in range(nslots): ld.acquire()
    dt=time.perf_counter()-t
    print(threads,'threads fill 64 rounds', f'{dt*1e6/nslots:.1f} us/round'); ld.stop()
