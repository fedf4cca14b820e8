import sys, zlib, numpy as np, torch
sys.path.insert(0, '.')
from tests import helpers as H
from tf2_b200 import capi, nets
from tf2_b200.network import NetWork, Runner
name, chw, spec, B = "flat_1x1_c1024_n256", (1024, 14, 14), dict(N=256, k=1), 2
rng = np.random.default_rng(zlib.crc32(name.encode()))
net = nets.chain(chw, [dict(N=chw[0], k=1, relu=1), dict(spec)], name)
x = H.random_input(rng, *chw, nonneg=True, B=B)
model = H.random_model(net, rng, x)
nw = NetWork(net, 0)
nw.InitFromCodes(model, None, max_images=B, variant=capi.VARIANT_MMA)
print(nw.layer_modes(B), nw.layer_modes(1), flush=True)
r = Runner(nw)
mode = sys.argv[1] if len(sys.argv) > 1 else "1"
nw.set_graph(int(mode))
out = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
print("run B ok", flush=True)
for b in range(B):
    tens, accs = H.oracle_tensors(net, model, x[b])
    assert np.array_equal(out[b], tens[net.result_tensor()])
print("parity ok", flush=True)
o1 = r.run_device(torch.from_numpy(x[:1].copy()).cuda()).cpu().numpy()
print("run 1 ok", np.array_equal(o1[0], out[0]), flush=True)
g = r.dump_acc(1, B).cpu().numpy()
print("dump ok", np.array_equal(g[0], accs[1]) if B == 1 else True, flush=True)
o1 = r.run_device(torch.from_numpy(x[:1].copy()).cuda()).cpu().numpy()
print("run 1 again ok", np.array_equal(o1[0], out[0]), flush=True)
