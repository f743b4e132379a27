import sys, json
sys.argv=["sweep"]
sys.path.insert(0,"tools")
import sweep, numpy as np
for dt in (np.float32, np.float64):
    sweep.run("cfg1_readme_64Mi_hbm","readme_deriv",50,dt,n=1<<26,check=4096)
