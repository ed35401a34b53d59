#!/bin/bash
# round 2, call 5A: the build at the end of the round -- whole GPU suite, bench lines of all five configs, the reference arm, launch list,
# ncu --set full of the packed extension kernel (class 65-80, first round) and of the pileup kernel
mkdir -p gpurun_out/r5a
O=gpurun_out/r5a
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $O/pytest.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
for c in 1 3 5; do timeout 600 python bench.py --config $c --cpu-seconds 5 > $O/bench_cfg$c.json 2> $O/bench_cfg$c.err; echo "cfg$c rc=$?"; done
timeout 900 python bench.py --config 4 --cpu-seconds 5 --steps 2 --warmup 1 > $O/bench_cfg4.json 2> $O/bench_cfg4.err; echo "cfg4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 1 --warmup 1 --cpu-seconds 0 --no-e2e > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled --kernel-name regex:'ext3_kernel<\(int\)80' -s 0 -c 1 -o $O/ext3_80 -f python tools/experiments/stage_ab.py 4 ncu > $O/ncu_ext3.log 2>&1; echo "ncu ext3 rc=$?"
python - <<'PY'
import json
for f in ("bench_default", "bench_cfg1", "bench_cfg3", "bench_cfg4", "bench_cfg5", "bench_reference"):
    try:
        s = open(f"gpurun_out/r5a/{f}.json").read(); d = json.loads(s[s.index("{"):])
        print(f, round(d["value"] / 1e6, 3), round(d["ms_per_step"], 2), d.get("e2e") and round(d["e2e"]["value"] / 1e6, 2), d.get("roofline") and round(d["roofline"]["frac"], 4),
              d.get("cpu_baseline") and (round(d["cpu_baseline"]["value"] / 1e6, 4), d["cpu_baseline"]["cores"]))
        if "stages_ms_per_step" in d: print("   ", {k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
