"""A/B timing of several builds of libsolo_b200.so (run on a GPU box): python tools/gpu_ab.py n1,n2 libA.so libB.so ...
(builds are selected with SOLO_B200_LIB; put the candidates under tools/_ab/, which travels to the box)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ns = sys.argv[1]
for lib in sys.argv[2:]:
    env = dict(os.environ, SOLO_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_sweep.py"), ns, "auto"], env=env,
                         capture_output=True, text=True).stdout
    for l in out.splitlines():
        if l.startswith("solo"):
            print(os.path.basename(lib), l, flush=True)
