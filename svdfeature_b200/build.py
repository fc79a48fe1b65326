"""In-tree build of the native libraries (sm_100a only).

    libsvdgpu.so    -- CUDA kernels + the C ABI of include/svdgpu.h
    libsvdf_gpu.so  -- the C++ ISVDTrainer implementation (gpu_trainer.cpp) behind the
                       same C shim the reference is wrapped with (trainer_cabi.cpp)

nvcc cross-compiles without a GPU; the .so files are git-ignored but travel to
the GPU box with the tree.  ``python -m svdfeature_b200.build`` rebuilds.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_GPU = os.path.join(HERE, "libsvdgpu.so")
LIB_TRAINER = os.path.join(HERE, "libsvdf_gpu.so")
REFERENCE_ROOT = "/root/reference"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # the reference never fuses multiply-add (built -msse2); parity needs the same
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-Xptxas", "-v",
    "-shared", "-cudart", "shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*names):
    return [os.path.join(CSRC, n) for n in names]


GPU_UNITS = ["svdgpu_api.cu", "svdgpu_stream.cu", "svdgpu_mf.cu", "svdgpu_ordered.cu", "svdgpu_own.cu", "svdgpu_comm.cu", "svdgpu_ingest.cu", "svdgpu_pairs.cu", "svdgpu_svdpp.cu",
             "svdgpu_rank.cu"]
GPU_HEADERS = ["svdgpu_internal.h", "svdgpu_device.cuh", "svdgpu_fb.cuh", "svdgpu_scan.h", "svdgpu_ownplan.h"]


def build_gpu(force=False, verbose=False):
    """nvcc -c every unit in parallel (one process each), then link libsvdgpu.so."""
    hdrs = _sources(*GPU_HEADERS) + [os.path.join(ROOT, "include", "svdgpu.h")]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs, objs, logs = [], [], []
    for unit in GPU_UNITS:
        src = os.path.join(CSRC, unit)
        obj = os.path.join(objdir, unit.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            cmd = [NVCC] + [f for f in NVCC_FLAGS if f not in ("-shared",)] + ["-c", "-o", obj, src]
            if os.environ.get("SVDGPU_TUNE_BUILD"):  # fewer template instances: quick kernel iteration only
                cmd.insert(1, "-DSVDGPU_TUNE_BUILD")
            procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for unit, p in procs:
        log, _ = p.communicate()
        logs.append("==== %s\n%s" % (unit, log))
        if p.returncode != 0:
            sys.stderr.write(log)
            raise RuntimeError("nvcc failed for %s" % unit)
    if logs:
        with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
            f.write("\n".join(logs))
        if verbose:
            print("\n".join(logs))
    if procs or not os.path.exists(LIB_GPU):
        subprocess.check_call([NVCC, "-shared", "-cudart", "shared", "-o", LIB_GPU] + objs + ["-ldl"])
    return LIB_GPU


def build_trainer(force=False):
    src = _sources("gpu_trainer.cpp", "trainer_cabi.cpp")
    if not all(os.path.exists(s) for s in src):
        return None
    deps = src + _sources("apex_compat.h") + [os.path.join(ROOT, "include", "svdgpu.h")]
    if not force and not _newer(LIB_TRAINER, deps):
        return LIB_TRAINER
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall",
           "-I", os.path.join(ROOT, "include"), "-o", LIB_TRAINER] + src + [
        "-L", HERE, "-lsvdgpu", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return LIB_TRAINER


def build_all(force=False, verbose=False):
    build_gpu(force, verbose)
    build_trainer(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
