"""Plain numeric (numpy) Kinova Gen3 kinematics and inverse dynamics used to pin the oracle by containment
properties — the assertion form of the reference's own visual check (KPR/debug_script.m:94-123, which overlays
simulator/dynamics/rnea.m torques on the dumped bounds).  Constants: KPR/KinovaWithoutGripperInfo.h."""
import numpy as np

TRANS = np.array([[0, 0, 0.15643], [0, 0.005375, -0.12838], [0, -0.21038, -0.006375], [0, 0.006375, -0.21038],
                  [0, -0.20843, -0.006375], [0, 0.00017505, -0.10593], [0, -0.10593, -0.00017505], [0, 0, 0]])
ROLL = np.array([np.pi, np.pi / 2, -np.pi / 2, np.pi / 2, -np.pi / 2, np.pi / 2, -np.pi / 2])
MASS = np.array([1.3773, 1.1636, 1.1636, 0.9302, 0.6781, 0.6781, 0.5])
COM = np.array([[-0.000023, -0.010364, -0.07336], [-0.000044, -0.09958, -0.013278], [-0.000044, -0.006641, -0.117892],
                [-0.000018, -0.075478, -0.015006], [0.000001, -0.009432, -0.063883], [0.000001, -0.045483, -0.00965], [0.000281, 0.011402, -0.029798]])
INERTIA = np.array([
    [0.00457, 0.000001, 0.000002, 0.000001, 0.004831, 0.000448, 0.000002, 0.000448, 0.001409],
    [0.011088, 0.000005, 0, 0.000005, 0.001072, -0.000691, 0, -0.000691, 0.011255],
    [0.010932, 0, -0.000007, 0, 0.011127, 0.000606, -0.000007, 0.000606, 0.001043],
    [0.008147, -0.000001, 0, -0.000001, 0.000631, -0.0005, 0, -0.0005, 0.008316],
    [0.001596, 0, 0, 0, 0.001607, 0.000256, 0, 0.000256, 0.000399],
    [0.001641, 0, 0, 0, 0.00041, -0.000278, 0, -0.000278, 0.001641],
    [0.000587, 0.000003, 0.000003, 0.000003, 0.000369, -0.000118, 0.000003, -0.000118, 0.000609]]).reshape(7, 3, 3)
ARMATURE = np.array([8.03, 11.9962024615303644, 9.0025427861751517, 11.5806439316706360, 8.4665040917914123, 8.8537069373742430, 8.8587303664685315])
LINK_C = np.array([[0.0, -0.001297, -0.088375], [0.0, -0.0894, -0.007877], [0.0, -0.001502, -0.129375], [0.0, -0.08745, -0.013648],
                   [0.000001, -0.009023, -0.071752], [0.0, -0.041661, -0.009251], [0.0, -0.018585, -0.033462]])
LINK_G = np.array([[0.046358, 0.047354, 0.086], [0.046, 0.1354, 0.047501], [0.046, 0.047501, 0.127], [0.046, 0.13345, 0.042293],
                   [0.034999, 0.044023, 0.069252], [0.035, 0.076739, 0.044076], [0.0455, 0.056085, 0.030963]])
GRAVITY = 9.81


def rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def joint_rotation(i, q):
    return rot_x(ROLL[i]) @ rot_z(q)


def bezier(q0, qd0, qdd0, k_actual, t):
    """q, qd, qdd of the degree-5 Bezier trajectory (KPR/Trajectory.h:10-28), duration 1."""
    q = t ** 3 * (6 * t ** 2 - 15 * t + 10) * k_actual + q0 + qd0 * t - 6 * qd0 * t ** 3 + 8 * qd0 * t ** 4 - 3 * qd0 * t ** 5 \
        + (qdd0 * t ** 2) / 2 - (3 * qdd0 * t ** 3) / 2 + (3 * qdd0 * t ** 4) / 2 - (qdd0 * t ** 5) / 2
    qd = 30 * t ** 2 * (t - 1) ** 2 * k_actual + ((t - 1) ** 2 * (2 * qd0 + 4 * qd0 * t + 2 * qdd0 * t - 30 * qd0 * t ** 2 - 5 * qdd0 * t ** 2)) / 2
    qdd = 60 * t * (2 * t ** 2 - 3 * t + 1) * k_actual - (t - 1) * (qdd0 - 36 * qd0 * t - 8 * qdd0 * t + 60 * qd0 * t ** 2 + 10 * qdd0 * t ** 2)
    return q, qd, qdd


def forward_kinematics(q):
    """Frames of the 7 joints: list of (R, p) in the base frame (same chain as KPR/Dynamics.cu:69-81)."""
    R, p = np.eye(3), np.zeros(3)
    frames = []
    for i in range(7):
        p = p + R @ TRANS[i]
        R = R @ joint_rotation(i, q[i])
        frames.append((R.copy(), p.copy()))
    return frames


def rnea(q, qd, qda, qdda):
    """Numeric Newton-Euler in the reference's formulation (KPR/Dynamics.cu:83-181 with numbers instead of PZs;
    simulator/dynamics/rnea.m:129-243), plus armature * qdda.  qd: velocity, qda: auxiliary velocity."""
    z = np.array([0.0, 0.0, 1.0])
    w = np.zeros(3); wdot = np.zeros(3); w_aux = np.zeros(3)
    lin = np.array([0.0, 0.0, GRAVITY])
    F, N = [], []
    Rs = [joint_rotation(i, q[i]) for i in range(7)] + [np.eye(3)]
    for i in range(7):
        Rt = Rs[i].T
        lin = Rt @ (lin + np.cross(wdot, TRANS[i]) + np.cross(w, np.cross(w_aux, TRANS[i])))
        w = Rt @ w + qd[i] * z
        w_aux = Rt @ w_aux
        wdot = Rt @ wdot + np.cross(w_aux, qd[i] * z) + qdda[i] * z
        w_aux = w_aux + qda[i] * z
        F.append(MASS[i] * (lin + np.cross(wdot, COM[i]) + np.cross(w, np.cross(w_aux, COM[i]))))
        N.append(INERTIA[i] @ wdot + np.cross(w_aux, INERTIA[i] @ w))
    f = np.zeros(3); n = np.zeros(3)
    u = np.zeros(7)
    for i in range(6, -1, -1):
        n = N[i] + Rs[i + 1] @ n + np.cross(COM[i], F[i]) + np.cross(TRANS[i + 1], Rs[i + 1] @ f)
        f = Rs[i + 1] @ f + F[i]
        u[i] = n[2] + ARMATURE[i] * qdda[i]
    return u
