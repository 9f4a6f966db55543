"""Record the unmodified reference's checksums and conservation diagnostics for
bench.py's headline workload (JW baroclinic wave ne = 120, L = 30, strang,
dt = 33.333333 s) after every step: tests/golden/bench_checksums.json, which
bench.py's `parity` block compares the device state with.

    python tests/make_bench_checksums.py [nsteps [npatch]]

(npatch = 24, the decomposition of the multi-GPU bench lines, is stored under
"reference_npatch24": the reference's own result depends on the decomposition
at the level of its sensitivity to rounding)

About 35 GB of host memory and 45 minutes on one core (10 minutes of serial
set-up, 75 s per step), which is why the numbers are committed rather than
recomputed.  The `device_one_gpu` entries of the file are the device's own
checksums at one GPU (gpurun_out/*bench_n1.json), the anchor multi-GPU runs are
held to (1e-10).  Needs /root/reference and `make -C oracle`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refdump  # noqa: E402

KEY = "JW baroclinic wave ne=120 L30 np=4 strang dt=33.3333s"
OUT = os.path.join(refdump.GOLDEN, "bench_checksums.json")

if __name__ == "__main__":
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    npatch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    script = "copy:0,3;energy:e0,3;checksum:c0"
    for i in range(1, nsteps + 1):
        script += ";step:1;checksum:c%d" % i
    script += ";copy:0,3;energy:e%d,3" % nsteps
    d = refdump.run_ref_dump("/tmp/ref120_p%d.bin" % npatch, "jw", script,
                             ["--resolution", "120", "--levels", "30", "--dt", "33333333u",
                              "--nogeometry", "1"], npatch=npatch, timeout=4 * 3600)
    table = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            table = json.load(f)
    entry = table.setdefault(KEY, {})
    if npatch != 6:
        entry["reference_npatch%d" % npatch] = {
            str(i): d["c%d.checksum" % i].tolist() for i in range(nsteps + 1)}
        with open(OUT, "w") as f:
            json.dump(table, f, indent=1, sort_keys=True)
        sys.exit(0)
    entry["reference"] = {str(i): d["c%d.checksum" % i].tolist() for i in range(nsteps + 1)}
    entry["reference_energy"] = {"0": d["e0.energy"].tolist(),
                                 str(nsteps): d["e%d.energy" % nsteps].tolist()}
    entry["reference_source"] = (
        "unmodified reference, oracle/_ref/ref_dump --case jw --resolution 120 --levels 30 "
        "--dt 33333333u --npatch 6 (tests/make_bench_checksums.py)")
    entry.setdefault("u_bound", 1e-6)
    with open(OUT, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
