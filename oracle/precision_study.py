"""TEST INFRASTRUCTURE ONLY -- CPU emulation of candidate tensor-core arithmetics for the path
solver (operand rounding to bf16/fp16/tf32 and hi+lo splits, fp32 accumulate), teacher-forced
against the fp32 oracle: prints max/rms error of the 10*tanh logits per stage.  The table in
DESIGN.md section 4 comes from `python oracle/precision_study.py 6`."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import path_solver_oracle as po
import htsp_b200
from htsp_b200.weights import synthetic_state_dict

def rb(t): return t.to(torch.bfloat16).to(torch.float32)
def rtf32(t):
    i = t.view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF   # round to 10-bit mantissa (approx RN)
    return i.view(torch.float32)
def rh(t): return t.to(torch.float16).to(torch.float32)
def mm(a, b, mode):
    if mode == 'fp16': return rh(a) @ rh(b)
    if mode == 'fp16x3':
        ah, bh = rh(a), rh(b); al, bl = rh(a-ah), rh(b-bh)
        return ah@bh + (ah@bl + al@bh)
    if mode == 'fp16x2b':   # only b split
        ah = rh(a); bh = rh(b); bl = rh(b-bh)
        return ah@bh + ah@bl
    if mode == 'fp16x2a':
        ah = rh(a); al = rh(a-ah); bh = rh(b)
        return ah@bh + al@bh
    if mode == 'fp32': return a @ b
    if mode == 'bf16': return rb(a) @ rb(b)
    if mode == 'tf32': return rtf32(a.contiguous()) @ rtf32(b.contiguous())
    if mode == 'bf16x3':
        ah, bh = rb(a), rb(b); al, bl = rb(a-ah), rb(b-bh)
        return ah@bh + (ah@bl + al@bh)
    if mode == 'bf16x2a':  # split only a (activations), b (weights/keys) single bf16
        ah = rb(a); al = rb(a-ah); bh = rb(b)
        return ah@bh + al@bh
    raise ValueError(mode)

def lin(x, w, b, mode):
    y = mm(x, w.t(), mode)
    return y if b is None else y + b

def heads(t): return t.reshape(t.size(0), t.size(1), 8, -1).transpose(1, 2)

def encode(sd, x, cfg):
    h = F.linear(x, sd["encoder.init_projection_layer.weight"], sd["encoder.init_projection_layer.bias"])
    if cfg.get('store_act'): h = cfg['store_act'](h)
    L = po.n_layers_of(sd)
    for i in range(L):
        p = f"encoder.attn_layers.{i}."
        q = heads(lin(h, sd[p+"Wq.weight"], None, cfg['enc_gemm']))
        k = heads(lin(h, sd[p+"Wk.weight"], None, cfg['enc_gemm']))
        v = heads(lin(h, sd[p+"Wv.weight"], None, cfg['enc_gemm']))
        s = mm(q, k.transpose(2,3), cfg['enc_attn'])/4
        pr = F.softmax(s, dim=3)
        o = mm(pr, v, cfg['enc_attn']).transpose(1,2).reshape(h.shape)
        h = h + lin(o, sd[p+"multi_head_combine.weight"], sd[p+"multi_head_combine.bias"], cfg['enc_gemm'])
        h = po._bn_eval(h, sd, p+"norm1")
        f = torch.relu(lin(h, sd[p+"feed_forward.0.weight"], sd[p+"feed_forward.0.bias"], cfg['enc_gemm']))
        f = lin(f, sd[p+"feed_forward.2.weight"], sd[p+"feed_forward.2.bias"], cfg['enc_gemm'])
        h = po._bn_eval(h + f, sd, p+"norm2")
        if cfg.get('store_act'): h = cfg['store_act'](h)
    return h

def run(sd, x, src, tgt, first, forced_pi, cfg, ref_logits):
    Bp, N, _ = x.shape; G = first.numel(); H=128
    emb = encode(sd, x, cfg)
    r = cfg['reset']
    K = heads(lin(emb, sd["decoder.Wk.weight"], None, r))
    V = heads(lin(emb, sd["decoder.Wv.weight"], None, r))
    QL = lin(emb, sd["decoder.Wq_last.weight"], None, r)          # per-node table
    QF = lin(emb, sd["decoder.Wq_first.weight"], None, r)
    qfix = lin(emb.mean(1, keepdim=True), sd["decoder.Wq_graph.weight"], None, r) \
         + lin(emb[torch.arange(Bp), src][:,None], sd["decoder.Wq_source.weight"], None, r) \
         + lin(emb[torch.arange(Bp), tgt][:,None], sd["decoder.Wq_target.weight"], None, r)
    Wc, bc = sd["decoder.multi_head_combine.weight"], sd["decoder.multi_head_combine.bias"]
    if cfg.get('fold'):
        K2 = lin(emb, Wc.t(), None, r)   # emb @ Wc  -> [B,N,128]:  u = O . K2^T
        c0 = emb @ bc
    srcG = src.view(Bp,1).expand(Bp,G); tgtG = tgt.view(Bp,1).expand(Bp,G)
    mask = torch.zeros(Bp,G,N); count = torch.zeros(Bp,G,dtype=torch.long)
    def visit(a):
        nonlocal count
        mask.scatter_(2, a[...,None], -math.inf); count = count+1
        partner = torch.where(a==srcG, tgtG, torch.where(a==tgtG, srcG, a))
        mask.scatter_(2, partner[...,None], -math.inf); count = count + (partner!=a).long()
        return partner
    firstG = first.view(1,G).expand(Bp,G)
    cur = visit(firstG)
    qf = QF[:, first]    # [Bp,G,128]
    maxerr = 0.0; flips = 0; tot = 0; errs=[]
    for t in range(N-2):
        q = qfix + qf + QL.gather(1, cur[...,None].expand(Bp,G,H))
        qh = heads(q)
        s = mm(qh, K.transpose(2,3), cfg['qk'])/4 + mask[:,None]
        p = F.softmax(s, dim=3)
        o = mm(p, V, cfg['pv']).transpose(1,2).reshape(Bp,G,H)
        if cfg.get('fold'):
            u = (mm(o, K2.transpose(1,2), cfg['logit']) + c0[:,None,:]) / math.sqrt(H)
        else:
            f = lin(o, Wc, bc, cfg['comb'])
            u = mm(f, emb.transpose(1,2), cfg['logit'])/math.sqrt(H)
        z = 10*torch.tanh(u) + mask
        ref = ref_logits[t]
        fin = torch.isfinite(ref)
        e = (z-ref)[fin].abs()
        maxerr = max(maxerr, e.max().item()); errs.append(e.pow(2).mean().item())
        flips += (z.argmax(2) != ref.argmax(2)).sum().item(); tot += Bp*G
        a = forced_pi.gather(2, count[...,None]).squeeze(2)
        cur = visit(a)
    return maxerr, math.sqrt(sum(errs)/len(errs)), flips/tot

if __name__ == '__main__':
    L = int(sys.argv[1]) if len(sys.argv)>1 else 6
    sd = synthetic_state_dict(1234, L)
    torch.manual_seed(3)
    Bp,N,G = 2,200,16
    x = torch.rand(Bp,N,2)*0.9+0.05
    src = torch.zeros(Bp,dtype=torch.long); tgt = torch.full((Bp,),N-1)
    first = torch.randperm(N)[:G]
    with torch.no_grad():
        res = po.rollout(sd, x, src, tgt, first, capture_logits=True)
        base = dict(enc_gemm='fp32', enc_attn='fp32', reset='fp32', qk='fp32', pv='fp32', comb='fp32', logit='fp32')
        cfgs = {
          'all fp32 (hoisted order)': base,
          'all fp32 folded': dict(base, fold=True),
          'all bf16': {k: 'bf16' for k in base},
          'all fp16': {k: 'fp16' for k in base},
          'enc tf32 only': dict(base, enc_gemm='tf32', enc_attn='tf32'),
          'logit tf32 only': dict(base, logit='tf32'),
          'all bf16x3 fold': dict({k: 'bf16x3' for k in base}, fold=True),
          'all fp16x3 fold (mode x3)': dict({k: 'fp16x3' for k in base}, fold=True),
          'fp16x3 fold, qk fp16 (mode fast)': dict({k: 'fp16x3' for k in base}, qk='fp16', fold=True),
          'enc fp16x3 only': dict(base, enc_gemm='fp16x3', enc_attn='fp16x3'),
          'enc gemm fp16x3, attn fp16': dict(base, enc_gemm='fp16x3', enc_attn='fp16'),
          'qk fp16 only': dict(base, qk='fp16'),
          'pv fp16 only': dict(base, pv='fp16'),
          'pv fp16x2a (P split only)': dict(base, pv='fp16x2a'),
          'pv fp16x2b (V split only)': dict(base, pv='fp16x2b'),
          'logit fp16 only fold': dict(base, logit='fp16', fold=True),
          'comb bf16 only': dict(base, comb='bf16'),
          'reset bf16 only': dict(base, reset='bf16'),
        }
        for name, cfg in cfgs.items():
            mx, rms, fl = run(sd, x, src, tgt, first, res.pi, cfg, res.logits)
            print(f"{name:40s} max|dz|={mx:.2e} rms={rms:.2e} argmax-flip={fl:.4f}", flush=True)
