"""Host emulation of conv3x3_small_kernel (kernels_elem.cu): the same block / thread / halo / weight-chunk index arithmetic in numpy
against a direct 3x3 convolution with weights [Cout][kh][kw][Cin] (k_conv_weight_prep layout). Checks the formulas, not the CUDA code."""
import numpy as np
rng=np.random.default_rng(0)
def emu(x, w, bias, COUT):
    N,H,W,Cin = x.shape
    TW,TH,CC = 32,8,64
    tiles_w=(W+TW-1)//TW; tiles_h=(H+TH-1)//TH
    y=np.zeros((N,H,W,COUT),np.float32)
    wf=w.reshape(COUT, 9*Cin)
    for b in range(N*tiles_h*tiles_w):
        bb=b; tx=bb%tiles_w; bb//=tiles_w; ty=bb%tiles_h; n=bb//tiles_h
        x0=tx*TW; y0=ty*TH
        acc=np.zeros((256,COUT,2),np.float64)
        for c0 in range(0,Cin,CC):
            sx=np.zeros((340,CC),np.float32)
            for idx in range(340*8):
                pix=idx>>3; part=idx&7
                gy=y0-1+pix//(TW+2); gx=x0-1+pix%(TW+2)
                if 0<=gy<H and 0<=gx<W: sx[pix,part*8:part*8+8]=x[n,gy,gx,c0+part*8:c0+part*8+8]
            sw=np.zeros(COUT*9*CC,np.float32)
            for idx in range(COUT*9*CC):
                co=idx//(9*CC); r=idx-co*9*CC; tap=r//CC; c=r-tap*CC
                sw[idx]=wf[co, tap*Cin+c0+c]
            for tid in range(256):
                px=tid&31; py=tid>>5
                for tap in range(9):
                    p=(py+tap//3)*(TW+2)+px+tap%3
                    for part in range(8):
                        f=sx[p,part*8:part*8+8]
                        for co in range(COUT):
                            wv=sw[(co*9+tap)*CC+part*8:(co*9+tap)*CC+part*8+8]
                            acc[tid,co,0]+= (f[0::2]*wv[0::2]).sum(); acc[tid,co,1]+=(f[1::2]*wv[1::2]).sum()
        for tid in range(256):
            px=tid&31; py=tid>>5; gx=x0+px; gy=y0+py
            if gx<W and gy<H: y[n,gy,gx,:]=acc[tid,:,0]+acc[tid,:,1]+bias
    return y
def ref(x,w,bias):
    N,H,W,Cin=x.shape; COUT=w.shape[0]
    xp=np.pad(x,((0,0),(1,1),(1,1),(0,0)))
    y=np.zeros((N,H,W,COUT),np.float64)
    for kh in range(3):
        for kw in range(3):
            y+=np.einsum('nhwc,oc->nhwo', xp[:,kh:kh+H,kw:kw+W,:], w[:,kh,kw,:])
    return y+bias
for (N,H,W,Cin,COUT) in [(2,10,35,128,3),(1,16,16,64,4),(1,9,33,64,4)]:
    x=rng.standard_normal((N,H,W,Cin)).astype(np.float32); w=rng.standard_normal((COUT,3,3,Cin)).astype(np.float32); b=rng.standard_normal(COUT).astype(np.float32)
    e=emu(x,w,b,COUT); r=ref(x,w,b)
    print((N,H,W,Cin,COUT), 'max abs diff', np.abs(e-r).max())
