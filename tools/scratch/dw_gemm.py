import torch
torch.backends.cuda.matmul.allow_tf32 = True
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for (M,Co,Ci) in [(32*512*128,96,64),(32*512*128,128,96),(32*512*128,64,10),(32*128*128,256,128),(32*128*64,128,128),(32*512*32,128,64),(32*512*16,64,32)]:
    dy=torch.randn(M,Co,device='cuda'); x=torch.randn(M,Ci,device='cuda')
    ref=torch.mm(dy.t(),x)
    gb=(M*(Co+Ci)*4)/1e9
    base=t(lambda: torch.mm(dy.t(),x))
    line=f"M={M} Co={Co} Ci={Ci}: mm {base*1e3:.0f}us ({gb/base*1e3:.0f} GB/s)"
    for rows in (2048,4096,8192,16384):
        P=M//rows
        f=lambda: torch.bmm(dy.view(P,rows,Co).transpose(1,2), x.view(P,rows,Ci)).sum(0)
        tt=t(f); err=float((f()-ref).abs().max()/ref.abs().max())
        line+=f" | bmm{rows} {tt*1e3:.0f}us e={err:.0e}"
    f=lambda: torch.mm(x.t(),dy).t()
    line+=f" | xT@dy {t(f)*1e3:.0f}us"
    print(line,flush=True)
