import torch, time
dev='cuda'
def t(fn, it=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/it*1000
shapes={'gx fwd (2048x1024)@(1024x1024)':(2048,1024,1024,False,False),
        'hh fwd (2048x256)@(256x1024)':(2048,256,1024,False,False),
        'dWih (1024x2048)@(2048x1024) tn':(1024,2048,1024,True,False),
        'dWhh (256x2048)@(2048x1024) tn':(256,2048,1024,True,False),
        'dh (2048x1024)@(1024x256)':(2048,1024,256,False,False)}
for name,(M,K,N,ta,tb) in shapes.items():
    A=torch.randn(K,M,device=dev).t() if ta else torch.randn(M,K,device=dev)
    B=torch.randn(K,N,device=dev)
    C=torch.empty(M,N,device=dev)
    res=[]
    for tf in (False,True):
        torch.backends.cuda.matmul.allow_tf32=tf
        res.append(t(lambda: torch.mm(A,B,out=C)))
    torch.backends.cuda.matmul.allow_tf32=True
    def three():
        torch.mm(A,B,out=C); C.addmm_(A,B); C.addmm_(A,B)
    r3=t(three)
    # accuracy of 3xTF32
    Ah=(A.contiguous().view(torch.int32)&-8192).view(torch.float32).view(A.shape) if not ta else None
    print('%-36s fp32 %.1f us  tf32 %.1f us  3 calls %.1f us'%(name,res[0],res[1],r3))
# accuracy check
torch.manual_seed(0)
A=torch.randn(2048,1024,device=dev); B=torch.randn(1024,1024,device=dev)
ref=(A.double()@B.double())
torch.backends.cuda.matmul.allow_tf32=False
e32=((A@B).double()-ref).abs().max().item()
def split(x):
    hi=(x.view(torch.int32)&-8192).view(torch.float32); return hi, x-hi
Ah,Al=split(A); Bh,Bl=split(B)
torch.backends.cuda.matmul.allow_tf32=True
C=Ah@Bh; C.addmm_(Ah,Bl); C.addmm_(Al,Bh)
e3=(C.double()-ref).abs().max().item()
e1=((A@B).double()-ref).abs().max().item()
# small terms first
C2=Ah@Bl; C2.addmm_(Al,Bh); C2.addmm_(Ah,Bh)
e3b=(C2.double()-ref).abs().max().item()
print('max abs err: fp32 %.3e  1xtf32 %.3e  3xtf32 %.3e  3xtf32(small first) %.3e  |ref|max %.1f'%(e32,e1,e3,e3b,ref.abs().max().item()))
