"""Debug of train_step.check: per-tensor differences between the sharded and the whole-batch step."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import odcp_b200
from odcp_b200 import train_step as ts, synthetic, targets, dist as yh_dist
from odcp_b200.optim import SGD, reset_state
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
if os.environ.get("IEEE"):
    torch.backends.cudnn.conv.fp32_precision = "ieee"; torch.backends.cuda.matmul.fp32_precision = "ieee"
n_total = 4 * world
case = synthetic.make_case("cfg4_check", 2, n_total, 13, 13, 5, 20, 416, 416, seed=104)
x_all = torch.rand(n_total, 416, 416, 3, generator=torch.Generator().manual_seed(7)) * 255.0
torch.manual_seed(11)
model = ts.YOLOv2Net(nets=ts.SMALL_NETS, bn=False, head_mid=32).to(dev)
state0 = {k: v.clone() for k, v in model.state_dict().items()}
rec, off, (lo, hi) = yh_dist.shard_case(case.rec, case.gt_off, n_total, rank, world)
st = ts.ShardedTrainStep(model)
share = st.step(x_all[lo:hi].to(dev), targets.records_to_tensor(rec, dev), torch.from_numpy(off).to(dev), len(rec))
sharded = {k: v.clone() for k, v in model.state_dict().items()}
model.load_state_dict(state0); reset_state()
opt = SGD(model.parameters(), **st.hyper); opt.zero_grad()
loss1 = model.get_loss_compact(x_all.to(dev), targets.records_to_tensor(case.rec, dev), torch.from_numpy(case.gt_off).to(dev))
loss1.backward(); opt.step()
rows = []
for k, v in model.state_dict().items():
    d = (sharded[k] - v).abs().max().item(); upd = (v - state0[k]).abs().max().item()
    rows.append((d / max(upd, 1e-12), k, d, upd, tuple(v.shape)))
rows.sort(reverse=True)
if rank == 0:
    print("cudnn conv fp32_precision:", getattr(getattr(torch.backends.cudnn, "conv", None), "fp32_precision", "n/a"))
    for r in rows[:8]: print("%.3e %s d=%.3e upd=%.3e %s" % r)
if world > 1: dist.destroy_process_group()
