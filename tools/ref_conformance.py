#!/usr/bin/env python
"""The reference's own test programs on the plugin's types.

PETSc's test suite runs many of its Vec / Mat / KSP test programs a second time on the device back end (`-vec_type cuda`,
`-mat_type aijcusparse`; the /*TEST ... TEST*/ block at the end of each source file).  Those variants are the reference's statement
of what a device type has to get right.  This tool

  build   (build container only: needs /root/reference)  compiles a list of such programs FROM THE SOURCES WHERE THEY LIE into
          baseline/_ref/petsc/bin/reftests/ (git-ignored, travels with gpurun like the library) and writes the manifest
          baseline/_ref/petsc/reftests.json: one case per single-rank device variant, with the reference's arguments translated
          (cuda -> b200, aijcusparse -> aijb200) and the same arguments on the host types (standard / aij);
  run     (GPU box; never touches /root/reference)  runs every case on the host types and on the b200 types with the plugin loaded
          and compares the two outputs (lines that name the type are dropped, as the reference's own `filter: grep -v type` does;
          numbers are compared like petscdiff -j: text must agree, numbers to a relative 1e-6); writes a JSON + markdown report.

tests/test_ref_conformance_gpu.py asserts the cases listed in tests/ref_conformance_expected.json.  No reference SOURCE is copied.
"""
import argparse
import json
import os
import re
import shlex
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "baseline", "_ref", "petsc")
BINDIR = os.path.join(OUT, "bin", "reftests")
MANIFEST = os.path.join(OUT, "reftests.json")
PLUGIN = os.path.join(ROOT, "petsc_plugin", "libpetscb200plugin.so")
BLASDIR = "/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs"

# programs whose TEST blocks carry a cuda / aijcusparse variant (grep -l over the reference tree), single source file each
PROGRAMS = [
    "vec/vec/tests/ex4.c", "vec/vec/tests/ex21.c", "vec/vec/tests/ex22.c", "vec/vec/tests/ex23.c", "vec/vec/tests/ex24.c", "vec/vec/tests/ex27.c",
    "vec/vec/tests/ex28.c", "vec/vec/tests/ex31.c", "vec/vec/tests/ex34.c", "vec/vec/tests/ex37.c", "vec/vec/tests/ex38.c", "vec/vec/tests/ex43.c",
    "vec/vec/tests/ex44.c", "vec/vec/tests/ex52.c", "vec/vec/tests/ex53.c", "vec/vec/tests/ex54.c", "vec/vec/tests/ex60.c", "vec/vec/tests/ex61.c",
    "vec/vec/tests/ex63.c", "vec/vec/tests/ex64.c", "vec/vec/tests/ex66.c", "vec/vec/tutorials/ex1.c", "vec/vec/tutorials/ex44.c",
    "vec/is/sf/tests/ex2.c", "vec/is/sf/tests/ex8.c", "vec/is/sf/tests/ex22.c",
    "ksp/ksp/tutorials/ex1.c", "ksp/ksp/tutorials/ex4.c", "ksp/ksp/tutorials/ex7.c", "ksp/ksp/tutorials/ex46.c", "ksp/ksp/tutorials/ex52.c",
    "ksp/ksp/tutorials/ex71.c", "ksp/ksp/tests/ex43.c", "ksp/ksp/tests/ex50.c", "ksp/ksp/tests/ex60.c",
    "mat/tests/ex1.c", "mat/tests/ex2.c", "mat/tests/ex5.c", "mat/tests/ex18.c", "mat/tests/ex23.c", "mat/tests/ex28.c", "mat/tests/ex62.c", "mat/tests/ex69.c",
    "mat/tests/ex70.c", "mat/tests/ex102.c", "mat/tests/ex123.c", "mat/tests/ex125.c", "mat/tests/ex132.c", "mat/tests/ex176.c", "mat/tests/ex217.c",
    "mat/tests/ex236.c", "mat/tests/ex237.c", "mat/tests/ex250.c", "mat/tests/ex251.c", "mat/tests/ex254.c", "mat/tests/ex261.c",
]
# tutorials without a device variant of their own: the reference's single-rank test arguments (or the arguments of a multi-rank
# test run on one rank) with the type options appended; T = "-mat_type X -vec_type Y", D = "-dm_mat_type X -dm_vec_type Y"
GENERIC = {
    "ksp/ksp/tutorials/ex5.c": [("1", "-pc_type jacobi -ksp_monitor -ksp_gmres_cgs_refinement_type refine_always", "T")],
    "ksp/ksp/tutorials/ex9.c": [("1", "-t 2 -pc_type jacobi -ksp_monitor -ksp_type gmres -ksp_gmres_cgs_refinement_type refine_always -s2_ksp_type bcgs -s2_pc_type jacobi -s2_ksp_monitor", "T")],
    "ksp/ksp/tutorials/ex15.c": [("1", "-ksp_view -user_defined_pc -ksp_gmres_cgs_refinement_type refine_always", "T")],
    "ksp/ksp/tutorials/ex16.c": [("1", "-ntimes 4 -ksp_gmres_cgs_refinement_type refine_always", "T")],
    "ksp/ksp/tutorials/ex23.c": [("1", "-ksp_monitor -ksp_gmres_cgs_refinement_type refine_always", "T"), ("cg_ilu", "-n 400 -ksp_type cg -pc_type ilu -ksp_monitor", "T"),
                                   ("cg_icc", "-n 400 -ksp_type cg -pc_type icc -ksp_monitor", "T"), ("bicg", "-n 100 -ksp_type bicg -pc_type jacobi -ksp_monitor", "T")],
    "ksp/ksp/tutorials/ex25.c": [("1", "-pc_type mg -ksp_type fgmres -da_refine 2 -ksp_monitor -mg_levels_ksp_monitor -mg_levels_ksp_norm_type unpreconditioned -ksp_view -pc_mg_type full", "D")],
    "ksp/ksp/tutorials/ex29.c": [("3", "-ksp_view -da_refine 2 -pc_type mg -pc_mg_distinct_smoothup -mg_levels_up_pc_type jacobi", "D")],
    "ksp/ksp/tutorials/ex32.c": [("4", "-pc_type mg -pc_mg_levels 2 -ksp_monitor_true_residual -ksp_rtol 1.e-10 -ksp_type cg -mg_levels_pc_type sor -mg_levels_ksp_type richardson -mg_levels_ksp_max_it 2 -mg_coarse_pc_type svd -da_refine 4", "D")],
    "ksp/ksp/tutorials/ex34.c": [("1", "-pc_type mg -pc_mg_type full -ksp_type fgmres -ksp_monitor -pc_mg_levels 3 -mg_coarse_pc_factor_shift_type nonzero -ksp_view", "D")],
    "ksp/ksp/tutorials/ex45.c": [("jacobi", "-ksp_monitor -da_grid_x 21 -da_grid_y 21 -da_grid_z 21 -pc_type jacobi", "D"), ("ilu", "-ksp_monitor -da_grid_x 21 -da_grid_y 21 -da_grid_z 21 -pc_type ilu", "D"),
                                   ("2", "-ksp_monitor -da_grid_x 21 -da_grid_y 21 -da_grid_z 21 -pc_type mg -pc_mg_levels 3 -mg_levels_ksp_type richardson -mg_levels_ksp_max_it 1 -mg_levels_pc_type bjacobi", "D")],
    # callers one layer up (Newton / time stepping on DMDA grids: re-assembled Jacobians, coloring, matrix-free, multigrid, fieldsplit)
    "snes/tutorials/ex5.c": [("1", "-snes_monitor -ksp_monitor_short -da_grid_x 17 -da_grid_y 17", "D"), ("mg", "-snes_monitor -pc_type mg -da_refine 2 -snes_view", "D"),
                               ("fd_color_ilu", "-snes_monitor -snes_fd_color -pc_type ilu -da_grid_x 12 -da_grid_y 12", "D"), ("mf", "-snes_monitor -snes_mf -pc_type none -da_grid_x 12 -da_grid_y 12", "D")],
    "snes/tutorials/ex19.c": [("ilu", "-da_refine 2 -snes_monitor_short -pc_type ilu", "D"),
                                ("fieldsplit", "-da_refine 2 -snes_monitor_short -pc_type fieldsplit -pc_fieldsplit_block_size 4 -pc_fieldsplit_type additive", "D")],
    "ts/tutorials/ex3.c": [("1", "-ts_monitor -ts_max_steps 5 -nox", "T")],
    # found by sweeping the reference's tests on the mock device: in-place LU / MatPermute end in MatHeaderMerge (our ops, the other
    # matrix's data and a NULL spptr); PCJacobiGetDiagonal after the sub-class applied the PC
    "mat/tests/ex15.c": [("1", "", "T")],
    "mat/tests/ex68.c": [("1", "", "T")],
    "ksp/ksp/tests/ex4.c": [("1", "-ksp_monitor -m 5 -pc_type jacobi -ksp_gmres_cgs_refinement_type refine_always", "T")],
    "ksp/ksp/tutorials/ex50.c": [("tut_1", "-da_grid_x 4 -da_grid_y 4 -mat_view", "D"), ("1", "-pc_type mg -pc_mg_type full -ksp_type cg -ksp_monitor -da_refine 3 -mg_coarse_pc_type svd -ksp_view", "D")],
}
TYPE_OPTS = {("T", False): "-mat_type aij -vec_type standard", ("T", True): "-mat_type aijb200 -vec_type b200",
             ("D", False): "-dm_mat_type aij -dm_vec_type standard", ("D", True): "-dm_mat_type aijb200 -dm_vec_type b200"}

# requirements a case may have (everything else -- kokkos, hip, datafilespath, external packages, complex ... -- disqualifies it)
OK_REQUIRES = {"cuda", "double", "!complex", "!single", "defined(PETSC_USE_LOG)", "!defined(PETSC_USE_64BIT_INDICES)", "!defined(PETSC_HAVE_MPIUNI)"}
DEVICE_WORDS = re.compile(r"\b(cuda|aijcusparse|seqaijcusparse|mpiaijcusparse|seqcuda|mpicuda|cusparse)\b")


def parse_test_block(text):
    """-> list of dicts {suffix, args, nsize, requires} with testset inheritance; loops {{a b}} keep their first value."""
    m = re.search(r"/\*TEST(.*?)TEST\*/", text, re.S)
    if not m:
        return []
    cases, stack = [], []   # stack of (indent, dict) for testset / test scopes
    for raw in m.group(1).splitlines():
        if not raw.strip() or raw.strip().startswith("#"):
            continue
        indent = len(raw) - len(raw.lstrip())
        key, _, val = raw.strip().partition(":")
        key, val = key.strip(), val.strip()
        while stack and stack[-1][0] >= indent:
            scope = stack.pop()[1]
            if scope.get("_kind") == "test":
                cases.append(scope)
        if key in ("test", "testset", "build"):
            parent = stack[-1][1] if stack else {}
            d = {"_kind": key, "args": parent.get("args", "") if parent.get("_kind") == "testset" else "",
                 "requires": parent.get("requires", "") if parent.get("_kind") == "testset" else "",
                 "nsize": parent.get("nsize", "1") if parent.get("_kind") == "testset" else "1", "suffix": parent.get("suffix", "") if parent.get("_kind") == "testset" else ""}
            stack.append((indent, d))
        elif stack:
            d = stack[-1][1]
            if key == "args":
                d["args"] = (d["args"] + " " + val).strip()
            elif key == "requires":
                d["requires"] = (d["requires"] + " " + val).strip()
            elif key in ("nsize", "suffix"):
                d[key] = val
    while stack:
        scope = stack.pop()[1]
        if scope.get("_kind") == "test":
            cases.append(scope)
    return [e for c in cases for e in expand_loops(c)]


LOOP = re.compile(r"\{\{([^}]*)\}\s*(?:separate output|shared output)?\s*\}")


def expand_loops(case):
    """{{a b c}separate output} / {{a b}shared output} / {{a b}}: one case per value (cartesian over several loops)."""
    m = LOOP.search(case["args"])
    if not m:
        return [case]
    out = []
    vals = m.group(1).split()
    if len(vals) > 3:
        vals = [vals[0], vals[1], vals[-1]]   # long loops (block sizes 1..15): first, second, last
    for v in vals:
        c = dict(case)
        c["args"] = case["args"][:m.start()] + v + case["args"][m.end():]
        c["suffix"] = "%s+%s" % (case["suffix"], v)
        out.extend(expand_loops(c))
    return out


def translate(args, device):
    vec, mat = ("b200", "aijb200") if device else ("standard", "aij")
    a = args
    a = re.sub(r"\b(seq|mpi)?aijcusparse\b", mat, a)
    a = re.sub(r"\b(seq|mpi)?cuda\b", vec, a)
    a = re.sub(r"(solver_type\s+)cusparse\b", r"\g<1>" + ("b200" if device else "petsc"), a)
    a = re.sub(r"-mat_cusparse_\S+\s+\S+", "", a)
    a = a.replace("-random_type curand", "")
    return shlex.split(a)


def build():
    if not os.path.isdir(REF) or not os.path.exists(os.path.join(OUT, "lib", "libpetsc.so")):
        print("ref_conformance build: needs /root/reference and baseline/_ref/petsc (oracle/build_ref.sh)")
        return 0
    os.makedirs(BINDIR, exist_ok=True)
    inc = ["-I%s/include" % REF, "-I%s/include" % OUT]
    lnk = ["-L%s/lib" % OUT, "-lpetsc", "-Wl,-rpath,$ORIGIN/../../lib", "-Wl,-rpath," + BLASDIR, "-Wl,-rpath-link," + BLASDIR, "-Wl,--allow-shlib-undefined", "-lm"]
    manifest = []
    for rel in PROGRAMS + sorted(GENERIC):
        src = os.path.join(REF, "src", rel)
        if not os.path.exists(src):
            continue
        name = rel.replace("/", "_")[:-2]
        exe = os.path.join(BINDIR, name)
        if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
            p = subprocess.run(["/usr/bin/gcc", "-O1", "-w", "-o", exe, src] + inc + lnk, capture_output=True, text=True)
            if p.returncode:
                print("  (does not build against this PETSc: %s)" % rel)
                continue
        for suffix, args, kind in GENERIC.get(rel, []):
            manifest.append({"program": rel, "exe": os.path.relpath(exe, ROOT), "suffix": "generic_" + suffix, "ref_args": args,
                             "host_args": shlex.split(args + " " + TYPE_OPTS[(kind, False)]), "b200_args": shlex.split(args + " " + TYPE_OPTS[(kind, True)])})
        if rel in GENERIC:
            continue
        for c in parse_test_block(open(src, errors="replace").read()):
            req = set(c["requires"].split())
            if c["nsize"] != "1" or not DEVICE_WORDS.search(c["args"]) or not req <= OK_REQUIRES or "cuda" not in req:
                continue
            if re.search(r"device_context|-use_gpu_aware_mpi|kokkos|viennacl|hip\b|-sf_backend|-mat_cusparse_storage_format (ell|hyb)|cupm|-log_view|densecuda|sellcuda|PETSC_DIR|DATAFILESPATH|solver_type cuda", c["args"]):
                continue
            manifest.append({"program": rel, "exe": os.path.relpath(exe, ROOT), "suffix": c["suffix"], "ref_args": c["args"],
                             "host_args": translate(c["args"], False), "b200_args": translate(c["args"], True)})
    # keep the cases this host-only MPIUNI build of the reference can run at all (on the host types, here, no GPU needed)
    kept = []
    for c in manifest:
        rc, _ = run_case(c, False)
        if rc == 0:
            kept.append(c)
        else:
            print("  (does not run on the host types in this PETSc build: %s)" % case_id(c))
    manifest = kept
    json.dump(manifest, open(MANIFEST, "w"), indent=1)
    print("ref_conformance: %d cases from %d programs -> %s" % (len(manifest), len({m["program"] for m in manifest}), os.path.relpath(MANIFEST, ROOT)))
    return 0


NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?")


def normalise(out):
    """Drops what legitimately differs between two back ends: lines naming a type or a solver package, the option-table warnings,
    and the object views of -ksp_view / -mat_view (an 'XXX Object:' line and the indented block under it)."""
    keep, in_view = [], False
    for line in out.splitlines():
        low = line.lower()
        if re.match(r"^\s*(KSP|PC|Mat|Vec|IS|PetscSF|DM) Object:", line):
            in_view = True
            continue
        if in_view and (line.startswith(" ") or line.startswith("\t")):
            continue
        in_view = False
        if "type" in low or "package used" in low or "option left" in low or "unused database option" in low or "warning!" in low or "could be spelling" in low:
            continue
        # names of the back end (a test that prints its solver / type name): map both sides to one spelling
        line = re.sub(r"(?i)\b(seq|mpi)?aijb200\b", lambda m: (m.group(1) or "") + "aij", line)
        line = re.sub(r"(?i)\b(seq|mpi)b200\b", lambda m: m.group(1), line)
        line = re.sub(r"(?i)\b(b200|petsc)\b", "SOLVER", line)
        keep.append(line.rstrip())
    return keep


def same_output(a, b, rtol=1e-6, atol=1e-10):
    la, lb = normalise(a), normalise(b)
    if len(la) != len(lb):
        return False, "line counts differ (%d vs %d)" % (len(la), len(lb))
    for x, y in zip(la, lb):
        if x == y:
            continue
        if NUM.sub("#", x) != NUM.sub("#", y):
            return False, "text differs: %r vs %r" % (x[:120], y[:120])
        for u, v in zip(NUM.findall(x), NUM.findall(y)):
            fu, fv = float(u), float(v)
            if abs(fu - fv) > atol + rtol * max(abs(fu), abs(fv)):
                return False, "number differs: %s vs %s in %r" % (u, v, x[:120])
    return True, ""


MOCK = os.path.join(ROOT, "tests", "mock", "libb200mock.so")


def run_case(case, device, timeout=25, mock=False):
    """mock=True (build container, no GPU): the plugin's C-ABI calls bind to the host test double tests/mock/libb200mock.so, so
    the plugin's HOST LOGIC runs on the CPU (tests/test_plugin_logic_mock_cpu.py); never used on a GPU box."""
    env = dict(os.environ, LD_LIBRARY_PATH=BLASDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    if mock and device:
        env["LD_PRELOAD"] = MOCK
    args = case["b200_args"] if device else case["host_args"]
    cmd = [os.path.join(ROOT, case["exe"])] + args + ["-dll_append", PLUGIN]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
        return p.returncode, p.stdout + p.stderr
    except subprocess.TimeoutExpired:
        return -9, "timeout"


def error_summary(out):
    """The message lines of a PETSc error (the two after the 'Error Message' banner) + the innermost stack frames."""
    err = [ln.split("PETSC ERROR:", 1)[1].strip() for ln in out.splitlines() if "PETSC ERROR:" in ln]
    msg = [e for e in err if e and not e.startswith("---") and not e.startswith("See http") and not e.startswith("PETSc ") and not e.startswith("Configure") and "Option Table" not in e and "source: command line" not in e]
    return " | ".join(msg[:5])[:700] if msg else " | ".join(out.strip().splitlines()[-6:])[:700]


def case_id(c):
    return "%s:%s" % (c["program"], c["suffix"])


def run(out_json, out_md, only=None, host_only=False, mock=False):
    manifest = json.load(open(MANIFEST))
    rows = []
    for c in manifest:
        if only and case_id(c) not in only:
            continue
        rc_h, out_h = run_case(c, False)
        row = {"case": case_id(c), "ref_args": c["ref_args"], "host_rc": rc_h}
        if host_only:
            row["status"] = "host ok" if rc_h == 0 else "host run fails"
        elif rc_h != 0:
            row["status"] = "skipped: fails on the host types in this PETSc build"
        else:
            rc_d, out_d = run_case(c, True, mock=mock)
            ok, why = same_output(out_h, out_d) if rc_d == 0 else (False, "exit code %d: %s" % (rc_d, error_summary(out_d)))
            row.update(b200_rc=rc_d, status="pass" if ok else "FAIL", why=why)
        rows.append(row)
        print(row["case"], row["status"], row.get("why", "")[:200], flush=True)
        if out_json:   # partial results survive a timeout of the whole run
            json.dump({"partial": True, "rows": rows}, open(out_json, "w"), indent=1)
    doc = {"passed": sum(r["status"] == "pass" for r in rows), "failed": sum(r["status"] == "FAIL" for r in rows), "skipped": sum(r["status"].startswith("skipped") for r in rows), "rows": rows}
    if out_json:
        json.dump(doc, open(out_json, "w"), indent=1)
    if out_md:
        with open(out_md, "w") as f:
            f.write("| reference test (program:variant) | reference's arguments | on the b200 types vs on the host types |\n|---|---|---|\n")
            for r in rows:
                f.write("| `%s` | `%s` | %s%s |\n" % (r["case"], r["ref_args"], r["status"], (" — " + r["why"][:160]) if r.get("why") else ""))
    print("ref_conformance: %d pass, %d fail, %d skipped" % (doc["passed"], doc["failed"], doc["skipped"]))
    return doc


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["build", "run", "hostcheck", "mockrun"])
    ap.add_argument("--json", default="")
    ap.add_argument("--md", default="")
    a = ap.parse_args()
    if a.cmd == "build":
        sys.exit(build())
    run(a.json, a.md, host_only=(a.cmd == "hostcheck"), mock=(a.cmd == "mockrun"))
