"""ORACLE (test infrastructure, not product): literal restatement of trajPlanner::mpcPlanner's plan / trajectory
interface, one planner object per scenario, plain Python floats and lists in the reference's own statement order.

Only tests/ may import this module.  It follows, method by method,

  trajectory_planner/include/trajectory_planner/mpcPlanner.cpp   (abbreviated MP.cpp; header mpcPlanner.h:108-175)
    updateMaxVel / updateMaxAcc        MP.cpp:249-255
    updateCurrStates (2 and 3 args)    MP.cpp:257-274      updateFovParam   MP.cpp:276-297
    updatePath                         MP.cpp:307-314
    updateDynamicObstacles             MP.cpp:316-341      updatePredObstacles   MP.cpp:343-373
    solveTraj                          MP.cpp:375-541  (assembly: oracle/mpc_assembly.py; solver: the reference's OSQP binary)
    makePlan                           MP.cpp:543-569      makePlanWithPred      MP.cpp:571-661
    findClosestObstacle                MP.cpp:663-708      getIntentComb         MP.cpp:710-769
    getTrajectoryScore                 MP.cpp:771-778      getConsistencyScore   MP.cpp:780-800
    getDetourScore                     MP.cpp:802-813      getSafetyScore        MP.cpp:815-852
    evaluateTraj                       MP.cpp:854-887
    getXRef                            MP.cpp:968-981      getReferenceTraj      MP.cpp:1199-1231
    getTrajectory / getPos / getVel / getAcc / getRef      MP.cpp:1234-1327

Arithmetic is IEEE double through Python's `math` (glibc libm, what the reference's C++ calls resolve to on x86-64), in the
reference's operation order: `Eigen::Vector3d::norm()` is sqrt(x*x + y*y + z*z), `pow(v, 2)` is v*v, `std::accumulate`
sums left to right, `maxCoeff(&i)` starts from element 0 and replaces on a strict `>` (Eigen 3.3, the version ROS Noetic
ships), so a NaN in element 0 is never replaced and NaNs elsewhere never win.

Not mirrored (no arithmetic on the path): RViz publishers, the point-cloud clustering thread (disabled in the reference,
MP.cpp:189-194 — static obstacles are handed in through `updateStaticObstacles`), and the two wall-clock mechanisms (the
0.15 s cut-off between candidates, MP.cpp:613, and the OSQP time limit, MP.cpp:442-444 — SURVEY.md §8(c) pins).
"""
from __future__ import annotations

import math

import numpy as np

from . import mpc_assembly as MA

FORWARD, LEFT, RIGHT, STOP = 0, 1, 2, 3          # dynamic_predictor/include/dynamic_predictor/utils.h:15-20


def _norm3(a, b):
    """(a - b).norm() for Eigen::Vector3d."""
    dx = a[0] - b[0]; dy = a[1] - b[1]; dz = a[2] - b[2]
    return math.sqrt(dx * dx + dy * dy + dz * dz)


def _max_coeff(v):
    """Eigen 3.3 DenseBase::maxCoeff(&index): max_coeff_visitor, init with coeff 0, replace on `value > res`."""
    res = v[0]; idx = 0
    for i in range(1, len(v)):
        if v[i] > res:
            res = v[i]; idx = i
    return res, idx


class RefPlanner:
    def __init__(self, params: MA.MpcParams, solve_qp):
        """solve_qp(QpBatch) -> dict(x, status, iter, ...) for B = 1 (the reference's OSQP binary behind
        oracle/bindings.py in the tests)."""
        self.p = params
        self.solve_qp = solve_qp
        self.horizon_ = params.horizon
        self.ts_ = params.ts
        self.currPos_ = [0.0, 0.0, 0.0]; self.currVel_ = [0.0, 0.0, 0.0]; self.currYaw_ = 0.0
        self.numHalfSpace_ = 0; self.halfMax_ = None; self.halfMin_ = None
        self.firstTime_ = True; self.stateReceived_ = False
        self.inputTraj_ = []; self.trajHist_ = []; self.lastRefStartIdx_ = 0
        self.staticObstacles_ = []                                   # obclustering_->getStaticObstacles(): (centroid, size, yaw)
        self.dynamicObstaclesPos_ = []; self.dynamicObstaclesSize_ = []
        self.obPredPos_ = []; self.obPredSize_ = []; self.obIntentProb_ = []
        self.currentStatesSol_ = []; self.currentControlsSol_ = []; self.ref_ = []
        self.candidateStates_ = []; self.candidateControls_ = []
        self.trajScore_ = []; self.trajWeightedScore_ = []
        self.obIdx_ = -1
        # bookkeeping for the tests (not in the reference): what every QP of the last call was and returned
        self.lastQps = []

    # ---- MP.cpp:249-373 --------------------------------------------------------------------------------
    def updateMaxVel(self, v):
        self.p.max_vel = v

    def updateMaxAcc(self, a):
        self.p.max_acc = a

    def updateCurrStates(self, pos, vel, yaw=None):
        self.currPos_ = [float(v) for v in pos]; self.currVel_ = [float(v) for v in vel]
        self.trajHist_.append(list(self.currPos_))
        if yaw is None:
            self.numHalfSpace_ = 0                                   # MP.cpp:261
        else:
            self.currYaw_ = float(yaw)
            self.updateFovParam()
        self.stateReceived_ = True

    def updateFovParam(self):                                        # MP.cpp:276-297 (87/2 is integer division: 43)
        maxAngle = self.currYaw_ - (87 // 2) * math.pi / 180.0
        minAngle = self.currYaw_ + (87 // 2) * math.pi / 180.0
        a1 = math.sin(maxAngle); b1 = -math.cos(maxAngle); c1 = a1 * self.currPos_[0] + b1 * self.currPos_[1]
        a2 = math.sin(minAngle); b2 = -math.cos(minAngle); c2 = a2 * self.currPos_[0] + b2 * self.currPos_[1]
        self.halfMax_ = [a1, b1, c1]; self.halfMin_ = [a2, b2, c2]
        self.numHalfSpace_ = 2

    def updatePath(self, path, ts):                                  # MP.cpp:307-314
        self.ts_ = ts; self.p.ts = ts
        self.inputTraj_ = [[float(v) for v in q] for q in path]
        self.firstTime_ = True; self.stateReceived_ = False
        self.trajHist_ = []; self.lastRefStartIdx_ = 0

    def updateStaticObstacles(self, obs):
        self.staticObstacles_ = [(list(map(float, c)), list(map(float, s)), float(y)) for c, s, y in obs]

    def updateDynamicObstacles(self, obstaclesPos, obstaclesVel, obstaclesSize):     # MP.cpp:316-341
        self.dynamicObstaclesPos_ = [[list(map(float, pos)) for _ in range(self.horizon_)] for pos in obstaclesPos]
        self.dynamicObstaclesSize_ = [[list(map(float, sz)) for _ in range(self.horizon_)] for sz in obstaclesSize]

    def updatePredObstacles(self, predPos, predSize, intentProb):    # MP.cpp:343-373: [ob][intent][step][3], [ob][4]
        if len(predPos):
            self.dynamicObstaclesPos_ = [[list(map(float, predPos[i][0][0])) for _ in range(self.horizon_)] for i in range(len(predPos))]
            self.dynamicObstaclesSize_ = [[list(map(float, predSize[i][0][0])) for _ in range(self.horizon_)] for i in range(len(predPos))]
            self.obPredPos_ = [[[list(map(float, q)) for q in tr] for tr in ob] for ob in predPos]
            self.obPredSize_ = [[[list(map(float, q)) for q in tr] for tr in ob] for ob in predSize]
            self.obIntentProb_ = [list(map(float, q)) for q in intentProb]
        else:
            self.dynamicObstaclesPos_ = []; self.dynamicObstaclesSize_ = []
            self.obPredPos_ = []; self.obPredSize_ = []; self.obIntentProb_ = []

    # ---- MP.cpp:375-541 --------------------------------------------------------------------------------
    def solveTraj(self, staticObstacles, dynamicObstaclesPos, dynamicObstaclesSize, xRef):
        """Returns (ok, statesSol, controlsSol).  Assembly through oracle/mpc_assembly.py (MP.cpp:891-1197)."""
        if self.firstTime_:
            self.currentStatesSol_ = []; self.currentControlsSol_ = []             # MP.cpp:378-381
        p = self.p
        N = p.N
        oxyz, osize, yaw, is_dyn = MA.obstacle_param(p, staticObstacles, dynamicObstaclesPos, dynamicObstaclesSize)
        x0 = np.array([self.currPos_ + self.currVel_])
        xr = np.array([[r[0:3] for r in xRef]])
        # linearisation point: previous plan at the same stage, unshifted, else the current position (MP.cpp:1042-1051)
        lin = np.zeros((1, N, 3))
        for k in range(N):
            lin[0, k] = self.currentStatesSol_[k][0:3] if k < len(self.currentStatesSol_) else self.currPos_
        # warm start (MP.cpp:485-509): previous plan where it exists, zeros elsewhere; dual = 0
        warm = np.zeros((1, p.n))
        for i in range(N + 1):
            if (not self.firstTime_) and i < len(self.currentStatesSol_):
                warm[0, 8 * i: 8 * i + 8] = self.currentStatesSol_[i]
        for i in range(N):
            if (not self.firstTime_) and i < len(self.currentControlsSol_):
                warm[0, 8 * (N + 1) + 5 * i: 8 * (N + 1) + 5 * i + 5] = self.currentControlsSol_[i]
        hs = None
        if self.numHalfSpace_:
            hs = (np.array([self.halfMax_]), np.array([self.halfMin_]))
        qb = MA.assemble_batch(p, x0, xr, oxyz[None], osize[None], yaw[None], is_dyn, lin, warm, half_space=hs)
        out = self.solve_qp(qb)
        self.lastQps.append(dict(qp=qb, out=out, num_obs=oxyz.shape[1]))
        if int(out["exitflag"][0]) != 0:                                           # MP.cpp:514-518
            return False, [], []
        x = out["x"][0]
        statesSol = [[float(v) for v in x[8 * i: 8 * i + 8]] for i in range(N + 1)]
        controlsSol = [[float(v) for v in x[8 * (N + 1) + 5 * i: 8 * (N + 1) + 5 * i + 5]] for i in range(N)]
        return True, statesSol, controlsSol

    # ---- MP.cpp:543-661 --------------------------------------------------------------------------------
    def makePlan(self):
        self.lastQps = []
        if self.firstTime_:
            self.currentStatesSol_ = []; self.currentControlsSol_ = []; self.ref_ = []
        staticObstacles = list(self.staticObstacles_)
        dynamicObstaclesPos = self.dynamicObstaclesPos_; dynamicObstaclesSize = self.dynamicObstaclesSize_
        if self.firstTime_:
            staticObstacles = []; dynamicObstaclesPos = []; dynamicObstaclesSize = []
        xRef = self.getXRef()
        ok, st, ct = self.solveTraj(staticObstacles, dynamicObstaclesPos, dynamicObstaclesSize, xRef)
        if ok:
            self.currentStatesSol_ = st; self.currentControlsSol_ = ct
            self.firstTime_ = False; self.ref_ = xRef
        return ok

    def makePlanWithPred(self):
        self.lastQps = []
        candidateStatesTemp = []; candidateControlsTemp = []; trajScore = []; intentType = []
        if self.firstTime_:
            self.candidateStates_ = []; self.candidateControls_ = []; self.trajWeightedScore_ = []; self.trajScore_ = []
            self.currentStatesSol_ = []; self.currentControlsSol_ = []; self.ref_ = []
        if not self.firstTime_:
            staticObstacles = list(self.staticObstacles_)
            dynamicObstaclesPos = self.dynamicObstaclesPos_; dynamicObstaclesSize = self.dynamicObstaclesSize_
        else:
            staticObstacles = []; dynamicObstaclesPos = []; dynamicObstaclesSize = []
        xRef = self.getXRef()
        if len(self.obPredPos_) and not self.firstTime_:
            obIdx, obstaclesPosComb, obstaclesSizeComb = self.getIntentComb(xRef)
            for i in range(len(obstaclesPosComb)):
                ok, statesSol, controlsSol = self.solveTraj(staticObstacles, obstaclesPosComb[i], obstaclesSizeComb[i], xRef)
                if ok:
                    candidateStatesTemp.append(statesSol); candidateControlsTemp.append(controlsSol)
                    trajScore.append(self.getTrajectoryScore(statesSol, controlsSol, staticObstacles, obstaclesPosComb[i], obstaclesSizeComb[i], xRef))
                    intentType.append(i)
            self.candidateStates_ = candidateStatesTemp; self.candidateControls_ = candidateControlsTemp
            if len(self.candidateStates_):
                self.firstTime_ = False
                validTraj = True
                bestTrajIdx = self.evaluateTraj(trajScore, obIdx, intentType)
                self.bestTrajIdx_ = bestTrajIdx
                self.currentStatesSol_ = self.candidateStates_[bestTrajIdx]
                self.currentControlsSol_ = self.candidateControls_[bestTrajIdx]
                self.trajScore_ = trajScore
                self.ref_ = xRef
            else:
                validTraj = False
        else:
            self.candidateStates_ = []; self.candidateControls_ = []; self.trajWeightedScore_ = []; self.trajScore_ = []
            validTraj, st, ct = self.solveTraj(staticObstacles, dynamicObstaclesPos, dynamicObstaclesSize, xRef)
            if validTraj:
                self.currentStatesSol_ = st; self.currentControlsSol_ = ct
                self.firstTime_ = False; self.ref_ = xRef
        return validTraj

    # ---- MP.cpp:663-769 --------------------------------------------------------------------------------
    def findClosestObstacle(self, xRef=None):
        obIdx = -1
        minDist = math.inf
        if self.firstTime_ or len(self.currentStatesSol_) < 2:
            for i in range(len(self.dynamicObstaclesPos_)):
                dist = _norm3(self.currPos_, self.dynamicObstaclesPos_[i][0])
                if dist < minDist:
                    minDist = dist; obIdx = i
            return obIdx
        for i in range(len(self.dynamicObstaclesPos_)):
            dist = 0.0
            for j in range(len(self.currentStatesSol_) // 3):
                state = self.currentStatesSol_[0][0:3]
                nextState = self.currentStatesSol_[1][0:3]
                ob = self.dynamicObstaclesPos_[i][0]
                trajDirectionAngle = math.atan2(nextState[1] - state[1], nextState[0] - state[0])
                obsDirectionAngle = math.atan2(ob[1] - state[1], ob[0] - state[0])
                weight = math.exp(-j)
                d = _norm3(state, ob)
                a = 3.0
                dist += weight * d * (a - math.cos(trajDirectionAngle - obsDirectionAngle))
                if dist > minDist:
                    break
            if dist < minDist:
                minDist = dist; obIdx = i
        return obIdx

    def intentWeights(self, obIdx):
        """The six hypothesis weights in their ORIGINAL order (MP.cpp:722-727 and again :866-871)."""
        pr = self.obIntentProb_[obIdx]
        return [pr[STOP], pr[LEFT], pr[RIGHT], pr[FORWARD], max(pr[LEFT], pr[FORWARD]), max(pr[RIGHT], pr[FORWARD])]

    def getIntentComb(self, xRef=None):
        obIdx = self.findClosestObstacle(xRef)
        self.obIdx_ = obIdx
        w6 = self.intentWeights(obIdx)
        weight = sorted((w6[i], i) for i in range(6))                # std::sort on pair<double,int>: ascending, lexicographic
        PP, PS = self.obPredPos_[obIdx], self.obPredSize_[obIdx]
        posTemp = [[PP[STOP]], [PP[LEFT]], [PP[RIGHT]], [PP[FORWARD]], [PP[LEFT], PP[FORWARD]], [PP[RIGHT], PP[FORWARD]]]
        sizeTemp = [[PS[STOP]], [PS[LEFT]], [PS[RIGHT]], [PS[FORWARD]], [PS[LEFT], PS[FORWARD]], [PS[RIGHT], PS[FORWARD]]]
        self.sortedCombo_ = [weight[5 - i][1] for i in range(6)]     # bookkeeping: hypothesis id at each sorted position
        intentCombPos = [list(posTemp[weight[5 - i][1]]) for i in range(6)]
        intentCombSize = [list(sizeTemp[weight[5 - i][1]]) for i in range(6)]
        for i in range(6):
            for j in range(len(self.obPredPos_)):
                if j != self.obIdx_:
                    _, maxIntent = _max_coeff(self.obIntentProb_[j])
                    intentCombPos[i].append(self.obPredPos_[j][maxIntent])
                    intentCombSize[i].append(self.obPredSize_[j][maxIntent])
        return obIdx, intentCombPos, intentCombSize

    # ---- MP.cpp:771-887 --------------------------------------------------------------------------------
    def getTrajectoryScore(self, states, controls, staticObstacles, obstaclePos, obstacleSize, xRef):
        return [self.getConsistencyScore(states), self.getDetourScore(states, xRef),
                self.getSafetyScore(states, staticObstacles, obstaclePos, obstacleSize)]

    def getConsistencyScore(self, state):
        numConsistencyStep = 10
        if self.firstTime_ or len(self.currentStatesSol_) == 0 or len(state) == 0:
            return 0.0
        maxStep = min(numConsistencyStep, min(len(self.currentStatesSol_), len(state)))
        if maxStep == 0:
            return 0.0
        totalDist = 0.0
        for i in range(maxStep):
            totalDist += _norm3(self.currentStatesSol_[i][0:3], state[i][0:3])
        totalDist /= maxStep
        return max(totalDist, 0.1)

    def getDetourScore(self, state, ref):
        totalDist = 0.0
        for i in range(len(state)):
            totalDist += _norm3(ref[i][0:3], state[i][0:3])
        totalDist /= len(state)
        return max(totalDist, 0.1)

    def getSafetyScore(self, state, staticObstacles, obstaclePos, obstacleSize):
        p = self.p
        totalDist = 0.0
        for i in range(len(state)):
            dist = 0.0; totalWeight = 0.0
            pos = [state[i][0], state[i][1], 0.0]
            for j in range(len(obstaclePos)):
                if i >= len(obstaclePos[j]):
                    raise IndexError("getSafetyScore reads obstaclePos[j][i] for every state i (MP.cpp:826): the prediction is too short")
                ob = [obstaclePos[j][i][0], obstaclePos[j][i][1], 0.0]
                maxSize = math.sqrt(obstacleSize[j][i][0] * obstacleSize[j][i][0] + obstacleSize[j][i][1] * obstacleSize[j][i][1])
                d = _norm3(pos, ob)
                weight = 1 - math.tanh(math.atanh(0.5) / (p.dynamic_safety_dist + maxSize) * d)
                dist += d * weight
                totalWeight += weight
            for j in range(len(staticObstacles)):
                c, s, _ = staticObstacles[j]
                ob = [c[0], c[1], 0.0]
                maxSize = math.sqrt((s[0] / 2) * (s[0] / 2) + (s[1] / 2) * (s[1] / 2))
                d = _norm3(pos, ob)
                weight = 1 - math.tanh(math.atanh(0.5) / (p.static_safety_dist + maxSize) * d)
                dist += d * weight
                totalWeight += weight
            dist = dist / totalWeight if totalWeight != 0.0 else (math.nan if dist == 0.0 else math.copysign(math.inf, dist))
            totalDist += dist
        totalDist /= len(state)
        return totalDist

    def evaluateTraj(self, trajScore, obIdx, intentType):
        self.trajWeightedScore_ = []
        consistentScore = [s[0] for s in trajScore]; detourScore = [s[1] for s in trajScore]; safetyScore = [s[2] for s in trajScore]

        def avg(v):
            a = 0.0
            for t in v:
                a += t
            return a / len(v)

        def div(a, b):                                               # IEEE division (Python raises on x / 0.0)
            if b == 0.0:
                return math.nan if (a == 0.0 or a != a) else math.copysign(math.inf, a) * math.copysign(1.0, b)
            return a / b
        consistentAvg = avg(consistentScore); detourAvg = avg(detourScore); safetyAvg = avg(safetyScore)
        for i in range(len(consistentScore)):
            consistentScore[i] = div(consistentAvg, consistentScore[i])
            detourScore[i] = div(detourAvg, detourScore[i])
            safetyScore[i] = div(safetyScore[i], safetyAvg)
        weight = self.intentWeights(obIdx)
        weightedScore = []
        for i in range(len(consistentScore)):
            # weight(intentType[i]): intentType[i] is the SORTED position of the candidate, weight is in the original order
            weightedScore.append(weight[intentType[i]] * (1.0 * consistentScore[i] + 1.0 * detourScore[i] + 1.0 * safetyScore[i]))
            self.trajWeightedScore_.append(weightedScore[i])
        _, bestTrajIdx = _max_coeff(weightedScore)
        return bestTrajIdx

    # ---- MP.cpp:968-981, 1199-1231 ---------------------------------------------------------------------
    def getXRef(self):
        return [[r[0], r[1], r[2], 0.0, 0.0, 0.0, 0.0, 0.0] for r in self.getReferenceTraj()]

    def getReferenceTraj(self):
        if len(self.inputTraj_) == 0:
            return [list(self.currPos_) for _ in range(self.horizon_)]
        leastDist = 1.7976931348623157e308
        maxForwardTime = 3.0
        maxForwardIdx = int(maxForwardTime / self.ts_)
        startIdx = self.lastRefStartIdx_
        searchEnd = min(self.lastRefStartIdx_ + maxForwardIdx, len(self.inputTraj_))
        for i in range(self.lastRefStartIdx_, searchEnd):
            dist = _norm3(self.currPos_, self.inputTraj_[i])
            if dist < leastDist:
                leastDist = dist; startIdx = i
        self.lastRefStartIdx_ = startIdx
        ref = []
        for i in range(startIdx, startIdx + self.horizon_):
            ref.append(list(self.inputTraj_[i]) if i < len(self.inputTraj_) else list(self.inputTraj_[-1]))
        return ref

    # ---- MP.cpp:1234-1327 ------------------------------------------------------------------------------
    def getTrajectory(self):
        return [s[0:3] for s in self.currentStatesSol_]

    def _interp(self, seq, off, t):
        idx = math.floor(t / self.ts_)
        dt = t - idx * self.ts_                                      # with the UNclamped index, as the reference does
        idx = max(0, min(idx, len(seq) - 1))
        nextIdx = min(idx + 1, len(seq) - 1)
        a, b = seq[idx], seq[nextIdx]
        return [a[off + c] + (b[off + c] - a[off + c]) / self.ts_ * dt for c in range(3)]

    def getPos(self, t):
        return list(self.currPos_) if len(self.currentStatesSol_) == 0 else self._interp(self.currentStatesSol_, 0, t)

    def getVel(self, t):
        return [0.0, 0.0, 0.0] if len(self.currentStatesSol_) == 0 else self._interp(self.currentStatesSol_, 3, t)

    def getAcc(self, t):
        return [0.0, 0.0, 0.0] if len(self.currentControlsSol_) == 0 else self._interp(self.currentControlsSol_, 0, t)

    def getRef(self, t):
        return list(self.currPos_) if len(self.ref_) == 0 else self._interp(self.ref_, 0, t)

    def getTs(self):
        return self.ts_

    def getHorizon(self):
        return self.horizon_
