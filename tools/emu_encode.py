"""Development helper: run integration/_build/turing_b200_batched against the CPU build of libhvb (tests/host_build.py: the
library's own sources over the warp emulator and a fake CUDA runtime) and compare bitstream + reconstruction with
oracle/_ref/turing_ref --asm 0.  usage: python tools/emu_encode.py [--rebuild] WxH frames [encoder options...]"""
import hashlib
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
LIBDIR = ROOT / "gpurun_out" / "emu_lib"


def build_lib():
    import host_build
    LIBDIR.mkdir(parents=True, exist_ok=True)
    files = tuple(sorted(p.name for p in host_build.CSRC.glob("*.cu") if not p.name.endswith("_tma.cu")))  # (TMA has no host emulation)
    host_build.build(LIBDIR, files, cpp_files=("havoc_b200.cpp", "hvb_encoder.cpp"), soname="libhvb.so")


def main():
    args = sys.argv[1:]
    if args and args[0] == "--rebuild":
        args = args[1:]
        build_lib()
    elif not (LIBDIR / "libhvb.so").exists():
        build_lib()
    res, frames, opts = args[0], int(args[1]), args[2:]
    w, h = map(int, res.split("x"))
    from turingcodec_b200 import synth
    work = ROOT / "gpurun_out" / "emu_work"
    work.mkdir(parents=True, exist_ok=True)
    clip = work / f"clip_{w}x{h}_{frames}.yuv"
    if not clip.exists():
        with open(clip, "wb") as f:
            for i in range(frames):
                for p in synth.frame(i, w, h, 8):
                    f.write(p.tobytes())

    def enc(binary, tag, extra, env):
        bit, rec = work / f"{tag}.bit", work / f"{tag}.yuv"
        cmd = [str(binary), "encode", "--input-res", res, "--frame-rate", "24", "--frames", str(frames), "-o", str(bit), "--dump-pictures", str(rec),
               *extra, *opts, str(clip)]
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        dt = time.time() - t0
        if r.returncode != 0:
            print(r.stdout[-1500:], r.stderr[-3000:])
            raise SystemExit(f"{tag} failed rc={r.returncode}")
        print(tag, f"{dt:.1f}s", [l for l in r.stderr.splitlines() if "hvbenc" in l])
        return hashlib.md5(bit.read_bytes()).hexdigest()[:8], hashlib.md5(rec.read_bytes()).hexdigest()[:8], bit.stat().st_size

    want = enc(ROOT / "oracle/_ref/turing_ref", "ref", ["--asm", "0"], dict(os.environ))
    got = enc(ROOT / "integration/_build/turing_b200_batched", "batched", [], dict(os.environ, LD_LIBRARY_PATH=str(LIBDIR), HVB_STATS="1"))
    print("ref    ", want)
    print("batched", got)
    print("IDENTICAL" if want == got else "DIFFERENT")
    return 0 if want == got else 1


if __name__ == "__main__":
    sys.exit(main())
