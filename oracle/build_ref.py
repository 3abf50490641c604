"""TEST INFRASTRUCTURE ONLY. Builds oracle/_ref/ from the reference sources where they lie.

The reference vendors the muSSP solver as a zip (/root/reference/mussp-master-6cf61b8.zip). Its core
(muSSP-master/muSSP/{Graph,Node,Sink}.cpp, main.cpp) is plain C++14 with no dependencies, so it is compiled
directly with the flags of the reference's own CMake file (mot3d/solvers/wrappers/muSSP/CMakeLists.txt:7-8:
-O2, C++14). The zip is unpacked into a throw-away temp dir; only binaries land in oracle/_ref/:

  oracle/_ref/libmussp_ref.so   unmodified muSSP core + oracle/mussp_shim.cpp (extern "C" re-host of wrapper.cpp)
  oracle/_ref/muSSP             the stock CLI (main.cpp), kept for cross-checks

wrapper.cpp itself needs Boost.Python, which this image lacks -> its driver logic is re-hosted by the shim.
oracle/_ref/ is git-ignored (never in history) but travels to the GPU box with gpurun.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ZIP = "/root/reference/mussp-master-6cf61b8.zip"
OUT = os.path.join(HERE, "_ref")
CXXFLAGS = ["-O2", "-std=c++14", "-fPIC", "-w"]


def have_ref():
    return os.path.exists(os.path.join(OUT, "libmussp_ref.so"))


def build(force=False, verbose=True):
    """Returns True when oracle/_ref/libmussp_ref.so exists afterwards."""
    if have_ref() and not force:
        return True
    if not os.path.exists(REF_ZIP):
        if verbose:
            print("[oracle/build_ref] %s not present (GPU box?) - using prebuilt oracle/_ref if any" % REF_ZIP)
        return have_ref()
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="mussp_src_")
    try:
        with zipfile.ZipFile(REF_ZIP) as z:
            members = [m for m in z.namelist() if m.startswith("muSSP-master/muSSP/") and
                       m.endswith((".cpp", ".h"))]
            z.extractall(tmp, members)
        src = os.path.join(tmp, "muSSP-master", "muSSP")
        core = [os.path.join(src, f) for f in ("Graph.cpp", "Node.cpp", "Sink.cpp")]
        so = os.path.join(OUT, "libmussp_ref.so")
        cmd = ["g++"] + CXXFLAGS + ["-shared", "-I", src, os.path.join(HERE, "mussp_shim.cpp")] + core + ["-o", so]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        cli = os.path.join(OUT, "muSSP")
        cmd = ["g++", "-O2", "-std=c++14", "-w", "-I", src, os.path.join(src, "main.cpp")] + core + ["-o", cli]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return have_ref()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref ready" if ok else "oracle/_ref NOT built")
    sys.exit(0 if ok else 1)
