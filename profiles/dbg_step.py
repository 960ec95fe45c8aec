import sys, time, torch
sys.path.insert(0,'.')
from rl_on_manifold_b200 import _lib, projection, synthetic
dev=torch.device('cuda:0')
nj=int(sys.argv[1]); B=int(sys.argv[2]); mode=int(sys.argv[3])
p=_lib.default_params("iiwa",nj); p.basis_mode=mode
q,dq,s,alpha=synthetic.device_batch("iiwa",B,3,dev,nj,p)
st=torch.zeros(B,dtype=torch.uint8,device=dev)
torch.cuda.synchronize()
for it in range(3):
    t=time.time()
    ddq,so=projection.step("iiwa",q,dq,s,alpha,p,n_ctrl_joints=nj,status=st)
    torch.cuda.synchronize()
    print("nj",nj,"B",B,"mode",mode,"ok %.4fs"%(time.time()-t),"deferred",int(((st&32)!=0).sum()), flush=True)
