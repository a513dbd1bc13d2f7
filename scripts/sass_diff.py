"""Which kernels' machine code changed since a git revision?

    python scripts/sass_diff.py <rev> [file.cu ...]        (default: every csrc/*.cu that differs from <rev>)

Compiles the named sources at <rev> (into a temporary directory) and in the working tree with the project's nvcc flags and
compares the SASS of every kernel instruction by instruction (addresses stripped, the per-translation-unit hash in mangled names
ignored).  Used to check that adding an opt-in experiment leaves the hardware-validated default kernels untouched.  CPU only."""
import importlib.util
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = "pan-tilt-zoom-slam_b200/csrc"


def kernels(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_", "", m.group(1))
            res[cur] = []
        elif cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            res[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return res


def main():
    rev = sys.argv[1]
    spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "pan-tilt-zoom-slam_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    flags = [f for f in b.FLAGS if f not in ("-Xptxas", "-v")]
    files = sys.argv[2:]
    if not files:
        diff = subprocess.run(["git", "diff", "--name-only", rev, "--", CSRC, "include"], cwd=ROOT, capture_output=True, text=True).stdout.split()
        headers = any(f.endswith((".h", ".cuh")) for f in diff)
        files = sorted(f for f in os.listdir(os.path.join(ROOT, CSRC)) if f.endswith(".cu")) if headers else \
            [os.path.basename(f) for f in diff if f.endswith(".cu")]
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "archive", rev, CSRC, "include"], cwd=ROOT, capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        changed = 0
        for f in files:
            objs = []
            for base in (tmp, ROOT):
                obj = os.path.join(tmp, ("old_" if base == tmp else "new_") + f[:-3] + ".o")
                subprocess.run([b.NVCC] + flags + ["-c", os.path.join(base, CSRC, f), "-o", obj], check=True, capture_output=True)
                objs.append(kernels(obj))
            old, new = objs
            for name, code in sorted(old.items()):
                if "EmptyKernel" in name:
                    continue
                if name not in new:
                    print("%-16s gone or renamed: %s" % (f, name[:90]))
                elif new[name] != code:
                    changed += 1
                    print("%-16s CHANGED (%d -> %d instructions): %s" % (f, len(code), len(new[name]), name[:90]))
            for name in sorted(set(new) - set(old)):
                print("%-16s new: %s" % (f, name[:90]))
        print("kernels with different machine code: %d" % changed)


if __name__ == "__main__":
    main()
