"""Back-to-back loss steps without host synchronisation (what a training loop does), stage by stage, each stage checked
against the first step's results.  python scripts/stress_loss.py [steps]"""
import sys
import torch
sys.path.insert(0, ".")
import colpali_b200 as cb
from oracle import li_oracle as O

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda:0")
q, d, _ = O.cfg3_inputs()
q, d = q.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
mod = cb.ColbertLoss()


def stage(name, fn, n):
    outs = [fn() for _ in range(n)]
    torch.cuda.synchronize()
    ref = outs[0]
    bad = sum(1 for o in outs if not all(torch.equal(a, b) for a, b in zip(o, ref)))
    print(f"stage {name}: {n} steps, {bad} differ from the first", flush=True)
    return bad == 0


def fwd_nograd():
    with torch.no_grad():
        return (mod(q, d).clone(),)


def fwd_grad():
    return (mod(q, d).detach().clone(),)


def fwd_bwd():
    q.grad = None; d.grad = None
    loss = mod(q, d)
    loss.backward()
    return (loss.detach().clone(), q.grad.clone(), d.grad[:2].clone())


ok = True
ok &= stage("forward, no grad (max mode)", fwd_nograd, steps)
ok &= stage("forward with grad (argmax mode), no backward", fwd_grad, steps)
ok &= stage("forward + backward", fwd_bwd, steps)
ok &= stage("forward, no grad again", fwd_nograd, steps)
print("STRESS", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
