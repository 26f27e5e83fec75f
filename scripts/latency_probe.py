import sys, os, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from oracle import synth, cscan
from vlite_b200 import _capi
for n in (100_000, 500_000, 32_000_000):
    w=128
    c=_capi.Corpus(w, capacity_rows=n); c.fill_synthetic(0,0,n)
    host = synth.corpus_rows(0,0,min(n,500_000),w)
    for nq in (1,2):
        q = synth.uniform_queries(1,nq,w); q[0]=host[777]
        dq=torch.from_numpy(q).cuda(); ds=torch.empty((nq,10),dtype=torch.int32,device="cuda"); dr=torch.empty((nq,10),dtype=torch.int64,device="cuda")
        res={}
        for mode,(sp,st) in {"fused":(0,0),"separate":(0,-2)}.items():
            c.set_tuning(sp,st); c.set_timing(True)
            tt=[];ts=[]
            for i in range(25):
                c.search(dq,10,0,None,ds,dr); t,s_=c.last_timing()
                if i>=5: tt.append(t); ts.append(s_)
            torch.cuda.synchronize()
            res[mode]=(round(float(np.mean(tt))*1e3,1), round(float(np.mean(ts))*1e3,1), ds.cpu().numpy().copy(), dr.cpu().numpy().copy())
            if n<=500_000:
                es,er=cscan.search(q,host[:n],10,0)
                assert np.array_equal(res[mode][2].view(np.uint32),es) and np.array_equal(res[mode][3].view(np.uint64),er), mode
        assert np.array_equal(res["fused"][2],res["separate"][2]) and np.array_equal(res["fused"][3],res["separate"][3])
        print(n,nq,"fused total/scan us",res["fused"][:2],"separate",res["separate"][:2], flush=True)
        c.set_timing(False)
        # host-buffer call latency
        c.set_tuning(0,0)
        t0=time.perf_counter()
        for i in range(200): c.search(q,10,0)
        print("   host-buffer call us", round((time.perf_counter()-t0)/200*1e6,1))
    c.close()
