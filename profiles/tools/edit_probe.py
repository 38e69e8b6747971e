import sys, os, json, numpy as np, torch
sys.path.insert(0, "husl-numpy_b200")
from nphusl import _cuda as cu
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
rgb = torch.randint(0,256,(8,2160,3840,3),dtype=torch.uint8,device="cuda"); px=rgb.numel()//3
res={}
e=cu.Edit(hue=(250,290),invert=True,l=(0.5,0))
for name,dt,npdt in (("f64",torch.float64,np.float64),("f32",torch.float32,np.float32)):
    hsl=torch.empty(rgb.shape,dtype=dt,device="cuda"); cu._rgb_to_husl(rgb,out=hsl,dtype=npdt)
    o=torch.empty_like(hsl)
    ms=timed(lambda: cu.edit_husl(hsl, e)); res["edit_in_place_"+name]=(round(ms,4), round(px*6*hsl.element_size()/ms/1e6,0))
    ms=timed(lambda: cu.edit_husl(hsl, e, out=o)); res["edit_out_of_place_"+name]=(round(ms,4), round(px*6*hsl.element_size()/ms/1e6,0))
    del hsl,o
print(json.dumps(res))
