"""TEST INFRASTRUCTURE, run by hand: random-shape fuzzing of the library's kernels executed on the host from their source
(tests/_emu_backend.py) against torch / the torch restatements of tests/_host_sim.py.

    python tests/fuzz_kernels.py conv|ctc|elementwise [--seed N] [--cases K]

conv: w2l_conv1d_fwd / _dgrad / _dgrad_wt / _wgrad through their own C wrappers vs torch conv1d autograd (T down to 1, Cout down
to 1, partial channel chunks, asymmetric zero padding, dilation); ctc: both schedules vs nn.CTCLoss in fp64 (zero-length inputs
and targets, infeasible pairs, repeated labels, 2..128 classes, logits or log-probs); elementwise: BatchNorm / activation / halo /
mask forward + backward, layout kernels, depthwise conv, log_softmax, colsum vs the restatements.  Combine with the sanitizer builds
(tools/host_sanitizers.sh shows the environment) to look for out-of-bounds / misaligned accesses on odd shapes."""
import argparse
import os
import random
import sys

import torch
import torch.nn.functional as TF

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _emu_backend as E  # noqa: E402
import _host_sim as H  # noqa: E402
from oracle import w2l_oracle as O  # noqa: E402,F401
from test_kernel_emu_gemm import ACT_NONE, DT_BF16, DT_F32, _bf, _pack_w, make_desc, rel_l2  # noqa: E402,F401

BF = torch.bfloat16


def fuzz_conv(rnd, cases):
    n_fail = 0
    for it in range(cases):
        B = rnd.choice([1,1,2,3]); k = rnd.choice([1,1,2,3,5,7,11]); d = rnd.choice([1,1,1,2,3])
        Cin = rnd.choice([64,72,80,96,128,136,192]); Cout = rnd.choice([1,8,16,29,40,64,100,128,130,256,272])
        pl = rnd.randint(0, (k-1)*d); pr = rnd.randint(0, (k-1)*d)
        T = rnd.choice([1,2,5,17,64,127,128,129,200,300])
        T_out = T + pl + pr - d*(k-1)
        if T_out < 1: continue
        g = torch.Generator().manual_seed(it)
        x = _bf(torch.randn(B,T,Cin,generator=g)); w = _bf(torch.randn(Cout,Cin,k,generator=g)/(Cin*k)**0.5)
        cout_pad = max(64,(Cout+15)//16*16); ldy=(Cout+7)//8*8
        xr = x.transpose(1,2).clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
        y_ref = TF.conv1d(TF.pad(xr,(pl,pr)), wr, dilation=d)
        dy = _bf(torch.randn(B,Cout,T_out,generator=g)); y_ref.backward(dy)
        xc, wc = x.to(torch.bfloat16), _pack_w(w,cout_pad)
        tag = (B,T,Cin,Cout,k,d,pl,pr)
        try:
            desc = make_desc(B,T_out,Cin,Cout,cout_pad,k,d,T,-pl,T_out,0,ldy,DT_F32,ACT_NONE)
            y = torch.full((B,T_out,ldy), float('nan')); E.conv1d_fwd(xc,wc,desc,y)
            e1 = rel_l2(y[:,:,:Cout], y_ref.detach().transpose(1,2))
            dyc = torch.zeros(B,T_out,cout_pad,dtype=torch.bfloat16); dyc[:,:,:Cout]=dy.transpose(1,2).to(torch.bfloat16)
            desc3 = make_desc(B,T_out,Cin,Cout,cout_pad,k,d,T,-pl,T_out,0,cout_pad)
            cin_pad=(Cin+15)//16*16; wt=torch.full((k,cin_pad,cout_pad),9.0,dtype=torch.bfloat16); E.pack_wt(w.permute(2,0,1).contiguous(),wt,Cout,Cin)
            dx=torch.full((B,T,Cin),float('nan'),dtype=torch.bfloat16); E.conv1d_dgrad_wt(dyc,wt,desc3,dx)
            e2 = rel_l2(dx.float(), xr.grad.transpose(1,2))
            dx1=torch.full((B,T,Cin),float('nan'),dtype=torch.bfloat16); E.conv1d_dgrad(dyc,wc,desc3,dx1)
            e2b = rel_l2(dx1.float(), xr.grad.transpose(1,2))
            dw=torch.full((k,Cout,Cin),3.0); E.conv1d_wgrad(dyc,xc,desc3,dw)
            e3 = rel_l2(dw, wr.grad.permute(2,0,1))
            ok = e1<2e-5 and e2<8e-3 and e2b<8e-3 and e3<2e-5
        except Exception as ex:
            ok=False; e1=e2=e2b=e3=str(ex)[:200]
        if not ok:
            n_fail += 1; print("FAIL", tag, e1, e2, e2b, e3, flush=True)
    return n_fail


def fuzz_ctc(rnd, cases):
    nf=0
    for it in range(cases):
        N=rnd.choice([1,2,3,5]); C=rnd.choice([2,3,5,29,31,32,33,64,100,128]); T=rnd.choice([1,2,3,7,31,32,33,64,65,100,257]); S=rnd.choice([0,1,2,3,10,40,130])
        g=torch.Generator().manual_seed(it)
        fl = rnd.random()<0.5
        x=torch.randn(N,T,C,generator=g)*rnd.choice([0.5,1.5,5.0])
        lp = x if fl else torch.log_softmax(x,-1)
        tg=torch.randint(1,C,(N,max(S,1)),generator=g,dtype=torch.int32) if C>1 else torch.zeros(N,max(S,1),dtype=torch.int32)
        if S==0: tg=torch.zeros(N,0,dtype=torch.int32)
        if S>2: tg[:,1::2]=tg[:,0::2][:,:tg[:,1::2].shape[1]]
        il=torch.tensor([rnd.randint(0,T) for _ in range(N)],dtype=torch.int32); tl=torch.tensor([rnd.randint(0,S) for _ in range(N)],dtype=torch.int32)
        il[0]=T
        for n in range(N): tg[n,tl[n]:]=0
        tag=(N,T,S,C,fl,il.tolist(),tl.tolist())
        try:
            ref_in = torch.log_softmax(lp.double(),-1) if fl else lp.double()
            # torch reference per utterance (skip il==0 rows: torch needs >=1? handles 0)
            xr = ref_in.clone().requires_grad_(True)
            l = torch.nn.CTCLoss(blank=0,reduction='mean',zero_infinity=True)(xr.transpose(0,1), tg if S>0 else torch.zeros(N,1,dtype=torch.int32), il, tl)
            (gr,) = torch.autograd.grad(l, xr)
            if fl:
                x2=lp.double().clone().requires_grad_(True)
                l2=torch.nn.CTCLoss(blank=0,reduction='mean',zero_infinity=True)(torch.log_softmax(x2,-1).transpose(0,1), tg if S>0 else torch.zeros(N,1,dtype=torch.int32), il, tl)
                (gr,)=torch.autograd.grad(l2,x2)
            for serial in (False,True):
                loss,nll,grad=E.ctc_loss_raw(lp,tg,il,tl,from_logits=fl,serial=serial)
                ok = abs(loss.item()-l.item())<=1e-4*max(1,abs(l.item())) and not torch.isnan(grad).any()
                gm=gr.abs().max().item()+1e-12
                # rows with il==0: torch grad 0
                err=(grad.double()-gr).abs().max().item()
                ok = ok and err<=2e-3*gm+1e-9
                if not ok:
                    nf+=1; print("FAIL",tag,serial,loss.item(),l.item(),err,gm,flush=True)
        except Exception as ex:
            nf+=1; print("EXC",tag,str(ex)[:300],flush=True)
    return nf


def fuzz_elementwise(rnd, cases):
    def rel(a,b):
        a,b=a.double().flatten(),b.double().flatten(); return float((a-b).norm()/(b.norm()+1e-30))
    def close(a,b,tol,what,tag):
        if a is None and b is None: return 0
        bad = torch.isnan(a.float()).any() or a.shape!=b.shape or rel(a.float(),b.float())>tol
        if bad: print("FAIL",what,tag, None if a.shape!=b.shape else rel(a.float(),b.float()), flush=True)
        return int(bad)
    nf=0
    for it in range(cases):
        g=torch.Generator().manual_seed(it)
        B=rnd.choice([1,2,3]); T=rnd.choice([1,2,3,9,31,32,33,100]); C=rnd.choice([8,16,24,64,72,256,264,520])
        pl=rnd.randint(0,min(6,T-1)); pr=rnd.randint(0,min(6,T-1)); act=rnd.choice([0,1,2])
        z=(torch.randn(B,T,C,generator=g)*2).to(BF); res=torch.randn(B,T,C,generator=g).to(BF) if rnd.random()<0.4 else None
        scale=torch.rand(C,generator=g)+0.5; shift=torch.randn(C,generator=g); rsc=torch.rand(C,generator=g)+0.5; rsh=torch.randn(C,generator=g)
        lens=torch.tensor([rnd.randint(0,T) for _ in range(B)],dtype=torch.int32) if rnd.random()<0.5 else None
        tag=(B,T,C,pl,pr,act,res is not None,None if lens is None else lens.tolist())
        kw=dict(lens=lens,res=res,res_scale=rsc if res is not None else None,res_shift=rsh if res is not None else None)
        a=E.bn_act_pad(z,scale,shift,B,T,C,pl,pr,act,0.0,0,**kw); b=H.bn_act_pad(z,scale,shift,B,T,C,pl,pr,act,0.0,0,**kw)
        nf+=close(a,b,4e-3,"bn_act_pad",tag)
        st_a=E.bn_stats(z,C); st_b=H.bn_stats(z,C); nf+=close(st_a,st_b,1e-5,"bn_stats",tag)
        gam=torch.rand(C,generator=g)+0.5; bet=torch.randn(C,generator=g)
        rm1,rv1,rm2,rv2=torch.zeros(C),torch.ones(C),torch.zeros(C),torch.ones(C)
        fa=E.bn_finalize(st_a,B*T,C,gam,bet,None,1e-3,0.1,rm1,rv1); fb=H.bn_finalize(st_b,B*T,C,gam,bet,None,1e-3,0.1,rm2,rv2)
        nf+=close(fa,fb,1e-4,"bn_finalize",tag); nf+=close(rv1,rv2,1e-4,"running_var",tag)
        dyp=torch.randn(B,pl+T+pr,C,generator=g).to(BF)
        dz_rows=T+rnd.choice([0,0,3])
        da=E.bn_act_bwd(dyp,z,fb[0],fb[1],fb[2],fb[3],gam,B,T,C,pl,pr,act,0.0,0,want_g=True,dz_rows=dz_rows,**kw)
        db=H.bn_act_bwd(dyp,z,fb[0],fb[1],fb[2],fb[3],gam,B,T,C,pl,pr,act,0.0,0,want_g=True,dz_rows=dz_rows,**kw)
        if B*T>1: nf+=close(da[0],db[0],1e-2,"dz",tag)      # one BatchNorm row: both sides compute ~0 from cancelling terms, no relative error to speak of
        nf+=close(da[1],db[1],1e-3,"red",tag); nf+=close(da[2],db[2],4e-3,"g",tag)
        # layout
        F_=rnd.choice([5,8,64,80]); Tn=rnd.choice([3,17,64,201]); k=rnd.choice([1,3,11]); s=rnd.choice([1,2,3]); d=rnd.choice([1,2]); pad=rnd.randint(0,min(5,Tn-1)); mode=rnd.choice([0,1])
        x=torch.randn(B,F_,Tn,generator=g); rows=(Tn+2*pad-d*(k-1)-1)//s+1
        ln=torch.tensor([rnd.randint(0,Tn) for _ in range(B)],dtype=torch.int32) if rnd.random()<0.5 else None
        if rows>=1:
            tag2=(B,F_,Tn,rows,k,s,d,pad,mode,None if ln is None else ln.tolist())
            nf+=close(E.im2col_ncw(x,rows,k,s,d,pad,mode,ln),H.im2col_ncw(x,rows,k,s,d,pad,mode,ln),1e-6,"im2col_ncw",tag2)
        Cc=rnd.choice([8,64,72]); xt=torch.randn(B,Tn,Cc,generator=g).to(BF)
        if rows>=1:
            a=E.im2col_tm(xt,rows,k,s,d,pad); b=H.im2col_tm(xt,rows,k,s,d,pad); nf+=close(a,b,1e-6,"im2col_tm",(B,Tn,Cc,rows,k,s,d,pad))
            dc=torch.randn(B,rows,k*Cc,generator=g).to(BF)
            nf+=close(E.col2im_tm(dc,Tn,Cc,k,s,d,pad),H.col2im_tm(dc,Tn,Cc,k,s,d,pad),5e-3,"col2im_tm",(B,Tn,Cc,rows,k,s,d,pad))
            # depthwise
            w=torch.randn(k,Cc,generator=g)
            ol=torch.tensor([rnd.randint(0,rows) for _ in range(B)],dtype=torch.int32) if rnd.random()<0.5 else None
            tg3=(B,Tn,Cc,rows,k,s,d,pad,None if ol is None else ol.tolist())
            nf+=close(E.depthwise_fwd(xt,w,rows,k,s,d,pad,ol),H.depthwise_fwd(xt,w,rows,k,s,d,pad,ol),6e-3,"dw_fwd",tg3)
            dy=torch.randn(B,rows,Cc,generator=g).to(BF)
            nf+=close(E.depthwise_dgrad(dy,w,Tn,k,d,pad,ol,stride=s),H.depthwise_dgrad(dy,w,Tn,k,d,pad,ol,stride=s),6e-3,"dw_dgrad",tg3)
            nf+=close(E.depthwise_wgrad(dy,xt,k,s,d,pad,ol),H.depthwise_wgrad(dy,xt,k,s,d,pad,ol),1e-4,"dw_wgrad",tg3)
        nf+=close(E.tm_to_ncw(xt,Tn-1 if Tn>1 else 1,Cc-3,x_row_offset=1 if Tn>1 else 0),H.tm_to_ncw(xt,Tn-1 if Tn>1 else 1,Cc-3,x_row_offset=1 if Tn>1 else 0),1e-6,"tm_to_ncw",(B,Tn,Cc))
        Cl=rnd.choice([2,29,33,100]); ld=(Cl+7)//8*8; lg=torch.randn(B,Tn,ld,generator=g)*3
        for m in (0,1):
            nf+=close(E.log_softmax(lg,Cl,m),H.log_softmax(lg,Cl,m),1e-5,"log_softmax%d"%m,(B,Tn,Cl))
        gr=torch.randn(B,Tn,Cl,generator=g); lp=H.log_softmax(lg,Cl,0)
        nf+=close(E.log_softmax_bwd(gr,lp,64 if Cl<=64 else 112),H.log_softmax_bwd(gr,lp,64 if Cl<=64 else 112),6e-3,"lsm_bwd",(B,Tn,Cl))
        m=torch.randn(rnd.choice([1,7,300]),ld,generator=g).to(BF)
        nf+=close(E.colsum(m,Cl),H.colsum(m,Cl),1e-4,"colsum",(m.shape,Cl))
    return nf


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("family", choices=["conv", "ctc", "elementwise"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cases", type=int, default=50)
    a = ap.parse_args()
    failures = {"conv": fuzz_conv, "ctc": fuzz_ctc, "elementwise": fuzz_elementwise}[a.family](random.Random(a.seed), a.cases)
    print("done, failures:", failures)
    sys.exit(1 if failures else 0)
