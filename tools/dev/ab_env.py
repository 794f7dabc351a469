"""Developer A/B driver: run bench.py once per value of one environment switch and print value / ms_per_step / e2e.
usage: python tools/dev/ab_env.py SB_STEPS_PER_GRAPH 8 10 20 [-- extra bench.py args]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    args = sys.argv[1:]
    extra = []
    if "--" in args:
        i = args.index("--")
        args, extra = args[:i], args[i + 1:]
    name, values = args[0], args[1:]
    for v in values:
        env = dict(os.environ)
        env[name] = v
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline"] + extra
        out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.strip().splitlines()
        try:
            d = json.loads(out[-1])
            print(f"{name}={v}: value {d['value']:.0f}  ms/step {d['ms_per_step']:.5f}  e2e {d['e2e']['value']:.0f}", flush=True)
        except Exception as ex:  # noqa: BLE001
            print(f"{name}={v}: no JSON line ({ex})", flush=True)


if __name__ == "__main__":
    main()
