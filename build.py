#!/usr/bin/env python
"""In-tree build of the native core: CUDA kernels for sm_100a (nvcc) + C++ graph/executor/binding (g++),
linked into hetu_b200/_C*.so with ninja (incremental, parallel).  No JIT cache: the .so lives in the tree
so it travels with the repository snapshot to the GPU box.

    python build.py            # build everything
    python build.py --tests    # also build the standalone kernel harnesses under build/
"""
import os
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(ROOT, "build")

CUDA_SOURCES = [
    "csrc/kernels/gemm_sm100.cu",
    "csrc/kernels/generic.cu",
    "csrc/kernels/attention_sm100.cu",
    "csrc/kernels/norm.cu",
    "csrc/kernels/elementwise.cu",
    "csrc/kernels/embedding_loss.cu",
    "csrc/kernels/optim.cu",
    "csrc/kernels/moe.cu",
    "csrc/kernels/quant_fp8.cu",
    "csrc/kernels/quant_block.cu",
    "csrc/kernels/softmax.cu",
    "csrc/kernels/symm_comm.cu",
]
CXX_SOURCES = [
    "csrc/core/device.cc",
    "csrc/core/ds.cc",
    "csrc/graph/graph.cc",
    "csrc/graph/exec.cc",
    "csrc/graph/zero_fused.cc",
    "csrc/graph/tp_fused.cc",
    "csrc/graph/ops_basic.cc",
    "csrc/graph/native_generic.cc",
    "csrc/graph/ops_nn.cc",
    "csrc/graph/ops_comm.cc",
    "csrc/graph/ops_optim.cc",
    "csrc/planner/dp_core.cc",
    "csrc/v1/embedding_cache.cc",
    "csrc/v1/ps_server.cc",
    "csrc/v1/ps_net.cc",
    "csrc/v1/ps_scheduler.cc",
    "csrc/runtime/symm_mem.cc",
    "csrc/runtime/symm_vmm.cc",
    "csrc/runtime/memory_pool.cc",
    "csrc/runtime/bfc_pool.cc",
    "csrc/runtime/runtime.cc",
    "csrc/runtime/rpc_client.cc",
    "csrc/binding/module.cc",
]
NVCC_FLAGS = "-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"


def main():
    import torch
    from torch.utils import cpp_extension

    os.makedirs(BUILD, exist_ok=True)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    nvcc = os.path.join(cuda_home, "bin", "nvcc")
    ext_suffix = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(ROOT, "hetu_b200", "_C" + ext_suffix)
    inc = cpp_extension.include_paths() + [os.path.join(cuda_home, "include"), sysconfig.get_paths()["include"]]
    inc_flags = " ".join("-isystem " + p for p in inc)
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cxx_flags = (f"-O2 -std=c++17 -fPIC -fvisibility=hidden -D_GLIBCXX_USE_CXX11_ABI={abi} -DTORCH_EXTENSION_NAME=_C "
                 f"-DTORCH_API_INCLUDE_EXTENSION_H -DUSE_C10D_NCCL -DUSE_C10D_GLOO -Wno-deprecated-declarations {inc_flags}")
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    ld_flags = (f"-shared -L{torch_lib} -Wl,-rpath,{torch_lib} -L{cuda_home}/lib64 -Wl,-rpath,{cuda_home}/lib64 "
                "-ltorch -ltorch_cpu -ltorch_python -lc10 -ltorch_cuda -lc10_cuda -lcudart -lpthread")
    cuda_srcs = [s for s in CUDA_SOURCES if os.path.exists(os.path.join(ROOT, s))]
    cxx_srcs = [s for s in CXX_SOURCES if os.path.exists(os.path.join(ROOT, s))]

    lines = [
        f"nvcc = {nvcc}", f"nvflags = {NVCC_FLAGS}", f"cxxflags = {cxx_flags}", f"ldflags = {ld_flags}", "",
        "rule cuda", "  command = $nvcc $nvflags -MD -MF $out.d -c $in -o $out", "  depfile = $out.d", "  deps = gcc",
        "  description = NVCC $in", "",
        "rule cxx", "  command = g++ $cxxflags -MD -MF $out.d -c $in -o $out", "  depfile = $out.d", "  deps = gcc",
        "  description = CXX $in", "",
        "rule link", "  command = g++ $in $ldflags -o $out.tmp && mv -f $out.tmp $out", "  description = LINK $out", "",
        "rule cudaexe", "  command = $nvcc $nvflags $in -o $out $libs", "  description = NVCC-EXE $out", "",
    ]
    objs = []
    for s in cuda_srcs:
        o = os.path.join(BUILD, s.replace("/", "_") + ".o")
        lines.append(f"build {o}: cuda {os.path.join(ROOT, s)}")
        objs.append(o)
    for s in cxx_srcs:
        o = os.path.join(BUILD, s.replace("/", "_") + ".o")
        lines.append(f"build {o}: cxx {os.path.join(ROOT, s)}")
        objs.append(o)
    lines.append(f"build {target}: link {' '.join(objs)}")
    defaults = [target]
    if "--tests" in sys.argv:
        gt = os.path.join(BUILD, "gemm_test")
        lines.append(f"build {gt}: cudaexe {ROOT}/csrc/tests/gemm_test.cu {ROOT}/csrc/kernels/gemm_sm100.cu")
        lines.append("  libs = -lcublas")
        at = os.path.join(BUILD, "attn_test")
        lines.append(f"build {at}: cudaexe {ROOT}/csrc/tests/attn_test.cu {ROOT}/csrc/kernels/attention_sm100.cu")
        lines.append("  libs = ")
        defaults += [gt, at]
    lines.append("default " + " ".join(defaults))
    with open(os.path.join(BUILD, "build.ninja"), "w") as f:
        f.write("\n".join(lines) + "\n")
    jobs = os.environ.get("MAX_JOBS", str(os.cpu_count() or 4))
    r = subprocess.run(["ninja", "-C", BUILD, "-j", jobs], cwd=ROOT)
    if r.returncode != 0:
        raise SystemExit("native build failed")
    print("built", target)


if __name__ == "__main__":
    main()
