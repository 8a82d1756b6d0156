"""A/B timing of two builds of libsolo_b200.so (run on a GPU box): python tools/gpu_ab.py libA.so libB.so [n,...]"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
libs = sys.argv[1:3]
ns = sys.argv[3] if len(sys.argv) > 3 else "4096,65536"
for lib in libs:
    env = dict(os.environ, SOLO_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_sweep.py"), ns, "auto"], env=env,
                         capture_output=True, text=True).stdout
    for l in out.splitlines():
        if l.startswith("solo12"):
            print(os.path.basename(lib), l, flush=True)
