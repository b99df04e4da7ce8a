import json, sys
d = json.load(open(sys.argv[1]))
print("fps", round(d["value"], 3), "ms", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 3), "launches", d["gpu_launches"], "clocks", d.get("clocks"))
for k, v in d["kernels"].items():
    print(f'{k:24s} {v["achieved"]:9.1f} {v["unit"]:8s} frac {v["frac"]:.3f} ms/step {v["ms_per_step"]:8.2f} share {v["share_of_step"]:.3f} avg_us {v["avg_launch_us"]:.1f}')
full = len(sys.argv) > 2
for k, v in d.get("gemm_layers", {}).items():
    if full or (k.count('.') <= 1 and (k in ('vit', 'dpt') or k.startswith('vit.') or k.startswith('fusion'))):
        print(f'{k:28s} {v["tflops"]:8.1f} TF/s frac {v["frac"]:.3f} ms/step {v["ms_per_step"]:8.2f} n {v["launches_per_step"]:.0f}')
