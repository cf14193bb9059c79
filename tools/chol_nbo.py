import os, sys, subprocess
for nbo in (512, 768, 1024, 1536):
    env = dict(os.environ, TN_CHOL_NBO=str(nbo))
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "chol_one.py"), sys.argv[1] if len(sys.argv) > 1 else "32768"], env=env, capture_output=True, text=True)
    print(nbo, out.stdout.strip(), out.stderr.strip()[-200:])
