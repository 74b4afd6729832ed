"""Build libv4l_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the
repo snapshot to the GPU box).  `python -m vision4leg_b200.build [--force]`."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libv4l_b200.so")

NVCC_FLAGS = [
  "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
  "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "177",
]


def sources():
  return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
  if not os.path.exists(LIB):
    return True
  t = os.path.getmtime(LIB)
  deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
    glob.glob(os.path.join(ROOT, "include", "*.h")) + [os.path.abspath(__file__)]
  return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
  if not force and not _stale():
    return LIB
  nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
  objdir = os.path.join(HERE, "build")
  os.makedirs(objdir, exist_ok=True)
  objs, procs = [], []
  for src in sources():
    obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
    objs.append(obj)
    cmd = [nvcc, "-c", src, "-o", obj, "-I", os.path.join(ROOT, "include"), "-I", CSRC] + NVCC_FLAGS
    if verbose:
      cmd += ["-Xptxas", "-v"]
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
  failed = False
  for src, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0 or verbose:
      sys.stderr.write(out.decode())
    failed |= p.returncode != 0
  if failed:
    raise RuntimeError("nvcc failed")
  tmp = LIB + ".tmp.%d" % os.getpid()
  subprocess.check_call([nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
  os.replace(tmp, LIB)
  return LIB


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
