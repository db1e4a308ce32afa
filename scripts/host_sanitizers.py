"""Race and memory check of the hand-written kernels without a GPU: the FFT passes and the fused single-field
kernels, compiled for the host (tests/host_emul: one OS thread per CUDA thread, __syncthreads / __syncwarp as
block / warp barriers, cp.async as an immediate copy -- the earliest moment the data may land), run under
ThreadSanitizer (a missing barrier is a data race on the shared tile) and under AddressSanitizer + UBSan.

    python scripts/host_sanitizers.py            # prints one line per build: reports found

compute-sanitizer's racecheck / memcheck on the device remain the authority; this is what can run in a
container without a GPU.
"""
import ctypes
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def dump_programs(out):
    import test_host_emulation_fused_cpu as F
    import test_step_gpu as T
    from gopf_b200 import pf as gpf
    from gopf_b200 import synthetic
    from gopf_b200._lib import check, lib

    def dump(m, rank, dt, tag):
        need, d = ctypes.c_int64(0), ctypes.c_int(-1)
        check(lib().gopf_model_fused_program_image(m._h, rank, ctypes.c_double(dt), None, ctypes.c_int64(0), ctypes.byref(need), None))
        prog = ctypes.create_string_buffer(need.value)
        check(lib().gopf_model_fused_program_image(m._h, rank, ctypes.c_double(dt), prog, need, None, ctypes.byref(d)))
        der = F._image(lib().gopf_model_derived_image, m._h, d.value, tail=(None,))
        open(os.path.join(out, tag + ".prog"), "wb").write(prog.raw)
        open(os.path.join(out, tag + ".der"), "wb").write(der.raw)

    dump(F.ch_pair([16, 16, 16])[0][0], 3, 0.1, "ch3")          # fast form, 3-D
    dump(F.ch_pair([32, 32])[0][0], 2, 0.1, "ch2")              # fast form, 2-D
    dump(F._pfc_pair(T, [32, 32])[0][0], 2, 0.1, "pfc2")        # general program through the rolled interpreters
    n = 16 ** 3
    m = gpf.NewModel()
    m.AddField(gpf.NewField("conc", n, synthetic.cahn_hilliard_initial(n, 3)))
    m.AddScalar(gpf.NewScalar("m1", -1.0))
    m.RegisterFunction("NOISE", gpf.WhiteNoise(1e-3, seed=11).Generate)
    m.AddEquation("dconc/dt = LAP conc^3 + m1*LAP conc + NOISE")
    m.SetKSpaceNoise(True)
    dump(m, 3, 0.01, "kn3")                                      # k-space noise inside the fused kernel


def main():
    emu = os.path.join(ROOT, "tests", "host_emul")
    inc = ["-I", emu, "-I", os.path.join(ROOT, "gopf_b200", "csrc")]
    with tempfile.TemporaryDirectory() as tmp:
        dump_programs(tmp)
        worst = 0
        for name, flags in (("thread", ["-fsanitize=thread"]), ("address+undefined", ["-fsanitize=address,undefined"])):
            for unit, main_src, args in (("emul_fft.cpp", "sanitize_fft_main.cpp", []),
                                         ("emul_fused.cpp", "sanitize_fused_main.cpp", [tmp])):
                exe = os.path.join(tmp, unit.replace(".cpp", "_" + name.split("+")[0]))
                subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-w", "-pthread", "-DGOPF_KNOISE", *flags, *inc,
                                os.path.join(emu, unit), os.path.join(emu, main_src), "-o", exe], check=True)
                env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0")
                r = subprocess.run([exe, *args], env=env, capture_output=True, text=True)
                text = r.stdout + r.stderr
                reports = text.count("WARNING: ThreadSanitizer") + text.count("ERROR: AddressSanitizer") + text.count("runtime error:")
                runs = text.count(" rc 0")
                print(f"{unit:16s} -fsanitize={name:18s} kernels runs ok: {runs:3d}   reports: {reports}   exit {r.returncode}")
                worst = max(worst, reports, 1 if r.returncode else 0)
        # self-test: with the barriers compiled out the same run must be full of reports
        exe = os.path.join(tmp, "emul_fft_nobarriers")
        subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-w", "-pthread", "-DGOPF_EMUL_NO_BARRIERS", "-fsanitize=thread", *inc,
                        os.path.join(emu, "emul_fft.cpp"), os.path.join(emu, "sanitize_fft_main.cpp"), "-o", exe], check=True)
        r = subprocess.run([exe], env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0"), capture_output=True, text=True)
        mutated = (r.stdout + r.stderr).count("WARNING: ThreadSanitizer")
        print(f"{'emul_fft.cpp':16s} -fsanitize=thread, barriers removed (self-test): reports: {mutated} (must be > 0)")
        if mutated == 0:
            worst = max(worst, 1)
        return worst


if __name__ == "__main__":
    sys.exit(1 if main() else 0)
