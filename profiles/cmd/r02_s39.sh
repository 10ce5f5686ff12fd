# round 2, call 39: smoke(), reference arm on the GPU box (unmodified reference from oracle/_ref), rotation angle batch sweep
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r02_s39_ref.err | cut -c1-900
python - <<'PY'
import time, torch
import unet_research_b200 as U
from unet_research_b200 import synthetic
from unet_research_b200.canonical import build_canonical
dev = torch.device("cuda")
m, _ = build_canonical(dev)
x = synthetic.make_image(584, 565, seed=1234).to(dev)
fov = synthetic.make_fov_mask(584, 565).to(dev)
for ab in (4, 5, 8, 10, 12):
    ev = U.RotationEval(m, num_iterations=359, return_num=25, angle_batch=ab)
    ev.predict_step((x, None, fov), 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ev.predict_step((x, None, fov), 0)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"angle_batch {ab}: {dt * 1e3:.1f} ms per 359 angles = {359 / dt:.0f} passes/s", flush=True)
PY
