"""Writes tests/golden/kinova_without_gripper.txt: the Kinova Gen3 (no gripper) robot description in the text format
the reference's robust-controller MEX reads (KRC/robot_models.cpp:20-122; KRC/kinova_without_gripper.txt), generated
from the robot constants the planner path already uses (tests/numeric_model.py = KPR/KinovaWithoutGripperInfo.h):

  twist i    joint axis in the joint frame (all revolute about z)
  inertia i  m, I_bar = Ic - m c^ c^ (inertia about the joint origin, joint frame), m c^   (row-major 3x3 each)
  Xtree i    transpose of the fixed parent->joint rotation (row-major), then the joint origin in the parent frame
  CoM i      centre of mass in the joint frame
  transI     transmission (armature) inertias; friction / damping are zero in this model

Usage: python tests/golden/make_robot_model.py   (the file is committed; tests read it, nothing reads /root/reference)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numeric_model as nm


def hat(c):
    return np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])


def fmt(values):
    return "<" + " ".join("%.30f" % float(v) for v in values) + ">"


def main():
    out = ["numJoints <7>", "gravity " + fmt([0.0, 0.0, -nm.GRAVITY])]
    for i in range(7):
        out.append("twist %d <0 0 1 0 0 0>" % i)
    for i in range(7):
        ch = hat(nm.COM[i])
        mch = nm.MASS[i] * ch
        I_bar = nm.INERTIA[i] - mch @ ch
        out.append("inertia %d " % i + fmt([nm.MASS[i]] + list(I_bar.reshape(-1)) + list(mch.reshape(-1))))
    for i in range(7):
        R = nm.rot_x(nm.ROLL[i]).T
        out.append("Xtree %d " % i + fmt(list(R.reshape(-1)) + list(nm.TRANS[i])))
    out.append("parent <-1 0 1 2 3 4 5>")
    for i in range(7):
        out.append("CoM %d " % i + fmt(nm.COM[i]))
    out.append("transI <" + " ".join("%.17f" % a for a in nm.ARMATURE) + ">")
    out.append("friction <0.0 0.0 0.0 0.0 0.0 0.0 0.0>")
    out.append("damping <0.0 0.0 0.0 0.0 0.0 0.0 0.0>")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kinova_without_gripper.txt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
