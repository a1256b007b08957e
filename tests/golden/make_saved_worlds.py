"""Packs the reference's 100 saved random worlds (kinova_src/saved_worlds/random/scene_*.csv; format
kinova_scenarios/load_saved_world.m:4-13: row 1 start configuration, row 2 goal configuration, row 3 NaN, rows 4.. one box
obstacle each as cx, cy, cz, sx, sy, sz) into tests/golden/worlds_saved_random.npz.  Run here, where /root/reference exists;
the GPU box only sees the .npz.  Obstacles are stored the way uarmtd_planner.m:189 hands them to the planner: centre plus
the three axis-aligned generators diag(side / 2) (simulator/worlds/obstacles/box_obstacle_zonotope.m:21-26), 12 doubles each."""
import glob
import os

import numpy as np

SRC = "/root/reference/kinova_src/saved_worlds/random"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    names, start, goal, n_obs, obs = [], [], [], [], np.zeros((0, 12))
    for path in sorted(glob.glob(os.path.join(SRC, "scene_*.csv"))):
        M = np.genfromtxt(path, delimiter=",")
        names.append(os.path.basename(path)[:-4])
        start.append(M[0, :7])
        goal.append(M[1, :7])
        boxes = M[3:, :6]
        z = np.zeros((len(boxes), 12))
        z[:, 0:3] = boxes[:, 0:3]
        z[:, 3], z[:, 7], z[:, 11] = boxes[:, 3] / 2, boxes[:, 4] / 2, boxes[:, 5] / 2
        n_obs.append(len(boxes))
        obs = np.vstack([obs, z])
    np.savez_compressed(os.path.join(HERE, "worlds_saved_random.npz"), names=np.array(names), start=np.array(start), goal=np.array(goal),
                        n_obs=np.array(n_obs, dtype=np.int32), obstacles=obs)
    print(len(names), "worlds,", int(np.sum(n_obs)), "boxes,", min(n_obs), "-", max(n_obs), "per world")


if __name__ == "__main__":
    main()
