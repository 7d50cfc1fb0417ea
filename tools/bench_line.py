"""Compact summary of a bench.py JSON line read from stdin (argv[1] = optional tag)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    k = d.get("kernels", {})
    parts = [tag, f"S={d['config'].get('streams_per_gpu')}", f"gpus={d['n_gpus']}",
             f"ms/step={d['ms_per_step']:.4f}", f"value={d['value']:.0f}"]
    if "p99_block_latency_ms" in d:
        parts.append(f"p99={d['p99_block_latency_ms']:.4f}ms")
    for name, v in k.items():
        parts.append(f"{name.replace('sh_', '').replace('_kernel', '')}={v['us']:.1f}us/{v['GBps']:.0f}GBps")
    if "e2e" in d:
        parts.append(f"e2e={d['e2e']['value']:.0f}")
    if d.get("clocks"):
        parts.append(f"clk={d['clocks'].get('sm_mhz')}")
    print(" ".join(parts))
