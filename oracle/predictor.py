"""ORACLE (test infrastructure, not product): literal restatement of dynamicPredictor::predictor's intent probabilities and
per-intent trajectory predictions — the step immediately upstream of mpcPlanner, whose outputs are
`updatePredObstacles`' arguments (predPos[ob][4][numPred+1], predSize likewise, intentProb[ob](4)).

Only tests/ may import this module.  It follows

  dynamic_predictor/include/dynamic_predictor/dynamicPredictor.cpp   (abbreviated PRED.cpp)
    initParam                 PRED.cpp:14-117   (derived constants paramf/l/r/s)
    intentProb                PRED.cpp:197-226  genTransitionMatrix :229-262   genTransitionVector :264-281
    predTraj                  PRED.cpp:283-329  genPoints :331-351
    modelForward              PRED.cpp:353-404  modelTurning :406-491          modelStop :493-505
    genTraj                   PRED.cpp:507-541  positionCorrection :543-572

in plain Python floats with the reference's statement order: the sampling loops run on DOUBLE counters that accumulate
(`for (double i = minAngle; i < maxAngle; i += 0.1)`), states advance by repeated addition (x += dt * vx), the variance is the
two-pass sum of squared deviations from the mean.  The occupancy map is FREE SPACE here (`map_->isInflatedOccupied(p)` is
false everywhere): sampled trajectories are never cut short and positionCorrection keeps the mean — what the map does is
perception, outside SURVEY.md §8.  Intent order FORWARD, LEFT, RIGHT, STOP (dynamic_predictor/utils.h:15-20).
"""
from __future__ import annotations

import dataclasses
import math

FORWARD, LEFT, RIGHT, STOP = 0, 1, 2, 3


@dataclasses.dataclass
class PredictorParams:
    """autonomous_flight/cfg/mpc_navigation/predictor_param.yaml, through initParam (PRED.cpp:14-117)."""
    prediction_size: int = 30
    prediction_time_step: float = 0.1
    min_turning_time: float = 2.0
    max_turning_time: float = 3.0
    prediction_z_score: float = 0.674
    max_front_prob: float = 0.5
    front_angle_deg: float = 25.0
    stop_velocity_threshold: float = 0.1
    prob_scale_param: float = 5.0

    def derived(self):
        paraml = (1 - self.max_front_prob) / (3 * self.max_front_prob - 1)            # PRED.cpp:75-76
        paramr = paraml
        front = self.front_angle_deg * math.pi / 180                                   # :87
        paramf = math.sqrt(front * front / (-2 * math.log(paraml * (1 + math.sin(front)) - paraml)))   # :88
        params = math.atanh(0.5) / self.stop_velocity_threshold                         # :99
        return dict(numPred=self.prediction_size, dt=self.prediction_time_step, zScore=self.prediction_z_score,
                    minTurn=self.min_turning_time, maxTurn=self.max_turning_time, frontAngle=front, stopVel=self.stop_velocity_threshold,
                    pscale=self.prob_scale_param, paramf=paramf, paraml=paraml, paramr=paramr, params=params)


def gen_transition_vector(c, theta, r, scale):                       # PRED.cpp:264-281
    pf = scale[0] * (math.exp(-0.5 * (theta / c["paramf"]) * (theta / c["paramf"])) + c["paraml"])
    pl = scale[1] * (c["paraml"] * (1 + math.sin(theta)))
    pr = scale[2] * (c["paramr"] * (1 - math.sin(theta)))
    ps = (1 - math.tanh(c["params"] / scale[3] * r))
    s = pr + pl + pf
    pr = (1 - ps) * pr / s
    pl = (1 - ps) * pl / s
    pf = (1 - ps) * pf / s
    out = [0.0] * 4
    out[FORWARD] = pf; out[LEFT] = pl; out[RIGHT] = pr; out[STOP] = ps
    return out


def intent_prob(c, pos_hist, vel_hist):
    """intentProb, PRED.cpp:197-226, for ONE obstacle: pos_hist / vel_hist [numHist][3], index 0 = newest."""
    P = [1.0 / 4] * 4
    nh = len(pos_hist)
    # The reference loops `for (j = 2; j < numHist; ++j)` and reads posHist_[i][numHist-j-2], which is index -1 on the last pass:
    # undefined behaviour (whatever precedes the vector's storage).  That pass is left out here and in the device kernel; every
    # defined pass is restated as written.  (Deviation recorded in DESIGN.md.)
    for j in range(2, nh - 1):
        prevPos = pos_hist[nh - j - 1]
        currPos = pos_hist[nh - j - 2]; currVel = vel_hist[nh - j - 2]
        older = pos_hist[nh - j]
        prevAngle = math.atan2(prevPos[1] - older[1], prevPos[0] - older[0])
        currAngle = math.atan2(currPos[1] - prevPos[1], currPos[0] - prevPos[0])
        theta = currAngle - prevAngle                                # genTransitionMatrix, :229-262
        if theta > math.pi:
            theta = theta - 2 * math.pi
        elif theta <= -math.pi:
            theta = theta + 2 * math.pi
        r = math.sqrt(currVel[0] * currVel[0] + currVel[1] * currVel[1])
        cols = []
        for i in range(4):
            scale = [1.0] * 4
            scale[i] = c["pscale"]
            cols.append(gen_transition_vector(c, theta, r, scale))
        newP = [0.0] * 4
        for row in range(4):                                          # Eigen dense mat * vec: row sums left to right
            a = 0.0
            for i in range(4):
                a += cols[i][row] * P[i]
            newP[row] = a
        P = newP
    return P


def model_stop(c, pos, vel, size):                                    # PRED.cpp:493-505
    pts = []; sizes = []
    sz = list(size)
    v = math.sqrt(vel[0] * vel[0] + vel[1] * vel[1])
    for _ in range(c["numPred"] + 1):
        pts.append(list(pos)); sizes.append(list(sz))
        sz[0] += 2 * min(v, c["stopVel"]) * c["dt"]
        sz[1] += 2 * min(v, c["stopVel"]) * c["dt"]
    return [pts], sizes


def model_forward(c, pos, vel, size):                                 # PRED.cpp:353-404
    v = math.sqrt(vel[0] * vel[0] + vel[1] * vel[1])
    angleInit = math.atan2(vel[1], vel[0])
    minVel = v - v; maxVel = v + v
    minAngle = angleInit - c["frontAngle"]; maxAngle = angleInit + c["frontAngle"]
    pts = []
    i = minAngle
    while i < maxAngle:
        j = minVel
        while j < maxVel:
            st = [pos[0], pos[1], j * math.cos(i), j * math.sin(i)]
            tr = [list(pos)]
            for _ in range(c["numPred"]):
                st = [st[0] + c["dt"] * st[2], st[1] + c["dt"] * st[3], st[2], st[3]]      # model * currState, model = I + dt on (p, v)
                tr.append([st[0], st[1], pos[2]])
            pts.append(tr)
            j += 0.1
        i += 0.1
    return pts, [list(size) for _ in range(c["numPred"] + 1)]


def model_turning(c, intent, pos, vel, size):                         # PRED.cpp:406-491
    v = math.sqrt(vel[0] * vel[0] + vel[1] * vel[1])
    angleInit = math.atan2(vel[1], vel[0])
    minVel = v - v; maxVel = v + v
    if intent == LEFT:
        endMin = c["frontAngle"] + angleInit
        endMax = (math.pi - c["frontAngle"]) + angleInit
        minAngVel = (math.pi / 2) / c["maxTurn"]
        maxAngVel = (math.pi / 2) / c["minTurn"]
    else:
        endMin = -(math.pi - c["frontAngle"]) + angleInit
        endMax = -c["frontAngle"] + angleInit
        minAngVel = (-math.pi / 2) / c["minTurn"]
        maxAngVel = (-math.pi / 2) / c["maxTurn"]
    pts = []
    i = minVel
    while i < maxVel:
        j = minAngVel
        while j < maxAngVel:
            endAngle = endMin
            while endAngle < endMax:
                angle = angleInit
                st = [pos[0], pos[1], i * math.cos(angle), i * math.sin(angle)]
                tr = [list(pos)]
                for _ in range(c["numPred"]):
                    st = [st[0] + c["dt"] * st[2], st[1] + c["dt"] * st[3], st[2], st[3]]
                    tr.append([st[0], st[1], pos[2]])
                    angle += j * c["dt"]
                    if intent == LEFT:
                        angle = min(angle, endAngle)
                    elif intent == RIGHT:
                        angle = max(angle, endAngle)
                    sp = math.sqrt(st[2] * st[2] + st[3] * st[3])
                    st[2] = sp * math.cos(angle); st[3] = sp * math.sin(angle)
                pts.append(tr)
                endAngle += 0.2
            j += 0.2
        i += 0.2
    return pts, [list(size) for _ in range(c["numPred"] + 1)]


def gen_points(c, intent, pos, vel, size):                            # PRED.cpp:331-351
    v = math.sqrt(vel[0] * vel[0] + vel[1] * vel[1])
    if v <= c["stopVel"]:
        return model_stop(c, pos, vel, size)
    if intent == FORWARD:
        return model_forward(c, pos, vel, size)
    if intent in (LEFT, RIGHT):
        return model_turning(c, intent, pos, vel, size)
    return model_stop(c, pos, vel, size)


def gen_traj(c, pts, sizes):                                          # PRED.cpp:507-541 (free space: positionCorrection keeps the mean)
    pred = []
    for i in range(c["numPred"] + 1):
        sumx = 0.0; sumy = 0.0; counter = 0
        for tr in pts:
            if i < len(tr):
                sumx += tr[i][0]; sumy += tr[i][1]; counter += 1
        if not counter:
            break
        meanx = sumx / counter; meany = sumy / counter
        svx = 0.0; svy = 0.0
        for tr in pts:
            svx += (tr[i][0] - meanx) * (tr[i][0] - meanx)
            svy += (tr[i][1] - meany) * (tr[i][1] - meany)
        varx = svx / counter; vary = svy / counter
        pred.append([meanx, meany, pts[0][0][2]])
        sizes[i][0] += 2 * math.sqrt(varx) * c["zScore"]
        sizes[i][1] += 2 * math.sqrt(vary) * c["zScore"]
    return pred, sizes


def predict_obstacle(p: PredictorParams, pos_hist, vel_hist, size):
    """predict(), PRED.cpp:162-195, for ONE obstacle -> (predPos[4][numPred+1][3], predSize[4][numPred+1][3], intentProb[4],
    samples per intent)."""
    c = p.derived()
    prob = intent_prob(c, pos_hist, vel_hist)
    pos = [float(v) for v in pos_hist[0]]; vel = [float(v) for v in vel_hist[0]]; size = [float(v) for v in size]
    pp = []; ps = []; ns = []
    for intent in range(4):
        pts, sizes = gen_points(c, intent, pos, vel, size)
        a, b = gen_traj(c, pts, sizes)
        pp.append(a); ps.append(b); ns.append(len(pts))
    return pp, ps, prob, ns
