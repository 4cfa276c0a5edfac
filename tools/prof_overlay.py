import sys, os, torch
sys.path.insert(0, '/root/repo')
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import ctypes as C
import bench
import rga3_release_b200 as vit
from rga3_release_b200 import _lib
T,H,W=16,448,448
frames=bench.synthetic_frames(T,H,W,0).cuda()
layer=bench.prompt_layer()
ops=[vit.FrameOp(mode=_lib.FRAME_LAYER, sx=i-8, sy=i-8) for i in range(T)]
spec=vit.OverlaySpec.from_rgba(layer, ops, device='cuda')
out=torch.empty((8192,1176),dtype=torch.bfloat16,device='cuda')
fc=_lib.Frames(frames.data_ptr(),T,H,W)
oc=spec.to_c(T)
for _ in range(5):
    _lib.check(_lib.lib().b200vit_overlay_patchify(C.byref(fc), C.byref(oc), 14,2,2,out.data_ptr(), torch.cuda.current_stream().cuda_stream),'p')
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    _lib.check(_lib.lib().b200vit_overlay_patchify(C.byref(fc), C.byref(oc), 14,2,2,out.data_ptr(), torch.cuda.current_stream().cuda_stream),'p')
e1.record(); torch.cuda.synchronize()
print('overlay_patchify us', e0.elapsed_time(e1)/20*1e3)
