import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_engine import _setup, _rel
from oracle import reference_port as rp
dev = torch.device("cuda:0")
d, st, eng, train_idx, B = _setup("cora", 2, dev)
for i in range(2):
    t = train_idx[i * B:(i + 1) * B]
    ref = rp.reference_step(st, t, apply_optim=True)
    noise = [h["noise"] for h in ref["hops"]]
    ref32 = rp.reference_step(st.fp32, t, gumbel_noise=noise, apply_optim=True)
    rec = eng.step(t.to(dev), gumbel_noise=[None if n is None else n.float().to(dev) for n in noise], apply_optim=True, record=True)
    for key, gk in (("gcn_c", "grads_c"), ("gcn_gf", "grads_gf"), ("gcn_z", "grads_z")):
        for name, g in ref[gk].items():
            print(f"step {i} grad {key}.{name}: gpu {_rel(rec['grads'][key][name], g):.2e}  cpu32 {_rel(ref32[gk][name], g):.2e}")
    for key, net, net32 in (("gcn_c", st.gcn_c, st.fp32.gcn_c), ("gcn_gf", st.gcn_gf, st.fp32.gcn_gf), ("gcn_z", st.gcn_z, st.fp32.gcn_z)):
        for (name, p), (_, p32) in zip(net.named_parameters(), net32.named_parameters()):
            got = eng.state_dicts()[key][name]
            print(f"step {i} param {key}.{name}: gpu {_rel(got, p.detach()):.2e}  cpu32 {_rel(p32.detach(), p.detach()):.2e}")
