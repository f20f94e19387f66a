"""torchrun --nproc-per-node N scripts/mgpu_check.py : ask-sharded sweep across N GPUs, replicated commit;
bindings must equal the oracle's (and therefore the single-GPU run's) on every rank."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from yunikorn_k8shim_b200 import Engine, synth, multigpu
from oracle import oracle_ctypes as oc
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ok_all = True
for snap, batch in ((synth.perf(900, 20, 100, masks=True), 256), (synth.hier(300, 3, 4, 2, 40, priorities=True), 128),
                    (synth.gangs(60, 40, 5, fill=1.4), 64), (synth.perf(masks=True), 4096),
                    (synth.perf(), 4096), (synth.perf(900, 20, 100), 64), (synth.poisoned_gangs(3), 16)):   # shared rows: fewer rows than ranks
    want = oc.run(snap)
    with Engine.for_snapshot(snap, batch=batch, device=lr, rank=rank, world=world) as e:
        mode = multigpu.attach(e, dist)
        torch.cuda.synchronize(); dist.barrier()
        t = time.time(); ask, node, _ = e.cycle(snap.n_asks); dt = time.time() - t
        st = e.stats()
        torch.cuda.synchronize(); dist.barrier()   # peers may still be signalling into this engine's sync block
    ok = np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"])
    agree = multigpu.check_agreement(dist, ask, node, device="cuda")
    ok_all = ok_all and ok and agree
    print(f"rank {rank}/{world} [{mode}] {snap.name}: identical_to_oracle={ok} replicas_agree={agree} cycle={dt*1e3:.1f} ms "
          f"evaluations(local)={st['evaluations']} sweep_ms={st['sweep_ms']:.2f}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
