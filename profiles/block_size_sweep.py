import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from thatsmyface_b200 import watermarking as W
n=64
imgs=torch.empty((n,bench.H,bench.W,3),dtype=torch.uint8,device='cuda'); bench.fill_images_device(imgs,0,17)
out=torch.empty_like(imgs)
for bs in ([int(a) for a in sys.argv[1:]] or (4,6,8,10,12,14,16)):
    wm=(torch.rand((bench.H//bs,bench.W//bs),device='cuda')<0.5).to(torch.uint8)*255
    for _ in range(2): W.embed_tensor(imgs,wm,0.1,bs,1,out=out)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): W.embed_tensor(imgs,wm,0.1,bs,1,out=out)
    e1.record(); torch.cuda.synchronize()
    ext=W.extract_tensor(out,imgs,0.1,bs,1)
    e2,e3=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(5): W.extract_tensor(out,imgs,0.1,bs,1)
    e3.record(); torch.cuda.synchronize()
    nat=torch.tensor([bench.image_kind(k)=="natural" for k in range(n)],device='cuda')
    ok=bool(torch.equal((ext[nat]>=128),(wm>=128).expand(int(nat.sum()),-1,-1)))
    print(f"bs {bs:2d}: embed {n*bench.PX*5/(e0.elapsed_time(e1)*1e-3)/1e6:9.0f} MP/s  extract {n*bench.PX*5/(e2.elapsed_time(e3)*1e-3)/1e6:9.0f} MP/s  bits recovered on natural images: {ok}")
