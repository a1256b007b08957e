"""Per-group busy/idle cycles of the task-scheduled reach kernel (profiling build).  Build the instrumented library first:
  cd armour-dev_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -DARMOUR_TASK_TIMING -c csrc/reach_kernels.cu -o /tmp/rk_tt.o \
    && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libarmour_b200_tt.so /tmp/rk_tt.o csrc/constraint_kernels.o csrc/controller_kernels.o csrc/armour_capi.o -lcudart
then run with ARMOUR_TUNE_TASKS=4 ARMOUR_TUNE_TASK_SCAP=1536 ARMOUR_TUNE_TASK_TCAP=256."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "armour-dev_b200"), os.path.join(ROOT, "tests")]
import armour_b200 as ab
ab.LIB_PATH = os.path.join(ab.PKG_DIR, "libarmour_b200_tt.so")
from problems import make_problem
p = ab.Planner(T=128, max_obstacles=20, device=0)
for s in range(2):
    q0, qd0, qdd0, _, obs = make_problem(100000 + s, 20)
    p.build(q0, qd0, qdd0, obs)
    print("reach_ms", p.last_build_ms()[1], flush=True)
