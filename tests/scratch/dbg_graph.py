import sys, copy, torch, torch.nn.functional as F
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from test_gpu_graphed import _make
from passt_b200.graphed import GraphedTrainStep
DEV='cuda'
B=4
torch.manual_seed(0)
waves=[0.1*torch.randn(B,320000,device=DEV) for _ in range(4)]
y=(torch.rand(B,527,device=DEV)<0.05).float()
mel_a, net_a = _make(1)
mel_b, net_b = copy.deepcopy(mel_a), copy.deepcopy(net_a)
params=lambda n:[p for k,p in n.named_parameters() if not k.startswith('head_dist')]
opt_a=torch.optim.AdamW(params(net_a), lr=1e-3, weight_decay=1e-4, fused=True)
opt_b=torch.optim.AdamW(params(net_b), lr=1e-3, weight_decay=1e-4, fused=True, capturable=True)
sd0=copy.deepcopy(net_a.state_dict())
step=GraphedTrainStep(mel_b, net_b, opt_b, F.binary_cross_entropy_with_logits, waves[0], y, warmup=2)
net_b.load_state_dict(sd0)
for st in opt_b.state.values():
    for k,v in st.items():
        if torch.is_tensor(v): v.zero_()
for i,w in enumerate(waves):
    torch.manual_seed(100+i)
    spec=mel_a(w).unsqueeze(1); logits,_=net_a(spec); la=F.binary_cross_entropy_with_logits(logits,y)
    opt_a.zero_grad(set_to_none=True); la.backward(); opt_a.step()
    torch.manual_seed(100+i); lb=step(w,y)
    print(i, float(la), float(lb))
for (k,pa),(_,pb) in list(zip(net_a.named_parameters(), net_b.named_parameters())):
    p0=sd0[k].to(DEV)
    da=(pa-p0).abs().max().item(); db=(pb-p0).abs().max().item(); dab=(pa-pb).abs().max().item()
    if k.startswith('blocks.1') or 'head' in k or 'norm.' in k or 'token' in k or 'pos' in k or 'patch' in k:
        print(f"{k:40s} |a-0| {da:.2e} |b-0| {db:.2e} |a-b| {dab:.2e}")
net_a.eval(); net_b.eval()
with torch.no_grad():
    xa=mel_a.eval()(waves[0]).unsqueeze(1)
    la, lb = net_a(xa)[0], net_b(xa)[0]
    print('eval logits a', la[0,:4].tolist(), 'b', lb[0,:4].tolist())
    net_b._wcache.clear()
    print('after cache clear b', net_b(xa)[0][0,:4].tolist())
    net_a._wcache.clear()
    print('after cache clear a', net_a(xa)[0][0,:4].tolist())
