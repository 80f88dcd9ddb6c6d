"""Stand-in for the `pybullet` module, used ONLY to run the unmodified reference
package (/root/reference/gym_pybullet_drones) in a container without pybullet.

TEST INFRASTRUCTURE — never imported by the product. It lets `oracle/gen_golden.py`
execute the reference's own `Physics.DYN` path, its force models and
`DSLPIDControl` to produce the golden vectors under `tests/golden/`.

Under `Physics.DYN` the reference uses Bullet only as
  (i)  a per-client body store (resetBasePositionAndOrientation / resetBaseVelocity,
       read back by getBasePositionAndOrientation / getBaseVelocity;
       reference BaseAviary.py:517-519, 862-872) and
  (ii) three closed-form converters (getMatrixFromQuaternion BaseAviary.py:836,
       getEulerFromQuaternion :518, getQuaternionFromEuler :488).
The closed forms below restate Bullet's published formulas
(b3Matrix3x3::setRotation, pybullet_getEulerFromQuaternion,
b3Quaternion::setEulerZYX + normalize; pybullet ^3.2.5, double precision build).
They cannot be checked against a real pybullet in this image ("parity
unpinned" against Bullet itself; pinned against the reference's Python).

`applyExternalForce/Torque` calls are recorded per client so the force models
(BaseAviary.py:715-811) can be read back by the golden generator.
"""
import math
import xml.etree.ElementTree as _et

DIRECT = 2
GUI = 1
LINK_FRAME = 1
WORLD_FRAME = 2
URDF_USE_INERTIA_FROM_FILE = 2
COV_ENABLE_RGB_BUFFER_PREVIEW = 0
COV_ENABLE_DEPTH_BUFFER_PREVIEW = 1
COV_ENABLE_SEGMENTATION_MARK_PREVIEW = 2
ER_TINY_RENDERER = 0
ER_SEGMENTATION_MASK_OBJECT_AND_LINKINDEX = 0
STATE_LOGGING_VIDEO_MP4 = 0

#: emulate Bullet's quaternion -> matrix -> quaternion round trip on read-back
ROUNDTRIP = False


class _Client:
    def __init__(self):
        self.bodies = {}
        self.next_id = 0
        self.forces = []   # (body, link, force3, pos3, flags)
        self.torques = []  # (body, link, torque3, flags)


_clients = {}
_next_client = [0]


def _c(physicsClientId=0):
    return _clients[physicsClientId]


def connect(mode, options=None):
    cid = _next_client[0]
    _next_client[0] += 1
    _clients[cid] = _Client()
    return cid


def disconnect(physicsClientId=0):
    _clients.pop(physicsClientId, None)


def resetSimulation(physicsClientId=0):
    c = _c(physicsClientId)
    c.bodies = {}
    c.next_id = 0
    c.forces = []
    c.torques = []


def setGravity(*a, **k): pass
def setRealTimeSimulation(*a, **k): pass
def setTimeStep(*a, **k): pass
def setAdditionalSearchPath(*a, **k): pass
def stepSimulation(*a, **k): pass
def configureDebugVisualizer(*a, **k): pass
def resetDebugVisualizerCamera(*a, **k): pass
def stopStateLogging(*a, **k): pass


def _link_offsets(path):
    """CoM offsets of the child links of a drone URDF (rotor links 0-3, CoM link 4)."""
    offs = []
    try:
        root = _et.parse(path).getroot()
    except Exception:
        return offs
    links = root.findall('link')
    for ln in links[1:]:
        org = ln.find('inertial/origin')
        xyz = [float(s) for s in org.attrib.get('xyz', '0 0 0').split()] if org is not None else [0., 0., 0.]
        offs.append(xyz)
    return offs


def loadURDF(fileName, basePosition=(0., 0., 0.), baseOrientation=(0., 0., 0., 1.),
             flags=0, physicsClientId=0, **kw):
    c = _c(physicsClientId)
    bid = c.next_id
    c.next_id += 1
    c.bodies[bid] = {
        'pos': tuple(float(v) for v in basePosition),
        'quat': tuple(float(v) for v in baseOrientation),
        'vel': (0., 0., 0.),
        'ang': (0., 0., 0.),
        'links': _link_offsets(fileName),
    }
    return bid


def getQuaternionFromEuler(rpy):
    roll, pitch, yaw = (float(v) for v in rpy)
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy = math.cos(hy), math.sin(hy)
    cp, sp = math.cos(hp), math.sin(hp)
    cr, sr = math.cos(hr), math.sin(hr)
    x = sr * cp * cy - cr * sp * sy
    y = cr * sp * cy + sr * cp * sy
    z = cr * cp * sy - sr * sp * cy
    w = cr * cp * cy + sr * sp * sy
    n = math.sqrt(x * x + y * y + z * z + w * w)
    return (x / n, y / n, z / n, w / n)


def getMatrixFromQuaternion(q):
    x, y, z, w = (float(v) for v in q)
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return (1.0 - (yy + zz), xy - wz, xz + wy,
            xy + wz, 1.0 - (xx + zz), yz - wx,
            xz - wy, yz + wx, 1.0 - (xx + yy))


def getEulerFromQuaternion(q):
    x, y, z, w = (float(v) for v in q)
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        return (0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y))
    if sarg >= 0.99999:
        return (0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y))
    return (math.atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz),
            math.asin(sarg),
            math.atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz))


def _mat_to_quat(m):
    e = [[m[0], m[1], m[2]], [m[3], m[4], m[5]], [m[6], m[7], m[8]]]
    tr = e[0][0] + e[1][1] + e[2][2]
    t = [0., 0., 0., 0.]
    if tr > 0.0:
        s = math.sqrt(tr + 1.0)
        t[3] = s * 0.5
        s = 0.5 / s
        t[0] = (e[2][1] - e[1][2]) * s
        t[1] = (e[0][2] - e[2][0]) * s
        t[2] = (e[1][0] - e[0][1]) * s
    else:
        i = (2 if e[1][1] < e[2][2] else 1) if e[0][0] < e[1][1] else (2 if e[0][0] < e[2][2] else 0)
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(e[i][i] - e[j][j] - e[k][k] + 1.0)
        t[i] = s * 0.5
        s = 0.5 / s
        t[3] = (e[k][j] - e[j][k]) * s
        t[j] = (e[j][i] + e[i][j]) * s
        t[k] = (e[k][i] + e[i][k]) * s
    return tuple(t)


def resetBasePositionAndOrientation(body, pos, quat, physicsClientId=0):
    b = _c(physicsClientId).bodies[body]
    b['pos'] = tuple(float(v) for v in pos)
    b['quat'] = tuple(float(v) for v in quat)


def resetBaseVelocity(body, linearVelocity=None, angularVelocity=None, physicsClientId=0):
    b = _c(physicsClientId).bodies[body]
    if linearVelocity is not None:
        b['vel'] = tuple(float(v) for v in linearVelocity)
    if angularVelocity is not None:
        b['ang'] = tuple(float(v) for v in angularVelocity)


def getBasePositionAndOrientation(body, physicsClientId=0):
    b = _c(physicsClientId).bodies[body]
    q = b['quat']
    if ROUNDTRIP:
        q = _mat_to_quat(getMatrixFromQuaternion(q))
    return b['pos'], q


def getBaseVelocity(body, physicsClientId=0):
    b = _c(physicsClientId).bodies[body]
    return b['vel'], b['ang']


def getLinkStates(body, linkIndices, computeLinkVelocity=0, computeForwardKinematics=0, physicsClientId=0):
    b = _c(physicsClientId).bodies[body]
    m = getMatrixFromQuaternion(b['quat'])
    out = []
    for li in linkIndices:
        o = b['links'][li]
        wp = (b['pos'][0] + m[0] * o[0] + m[1] * o[1] + m[2] * o[2],
              b['pos'][1] + m[3] * o[0] + m[4] * o[1] + m[5] * o[2],
              b['pos'][2] + m[6] * o[0] + m[7] * o[1] + m[8] * o[2])
        out.append((wp, b['quat'], tuple(o), (0., 0., 0., 1.), wp, b['quat'], b['vel'], b['ang']))
    return out


def applyExternalForce(objectUniqueId, linkIndex, forceObj, posObj, flags, physicsClientId=0):
    _c(physicsClientId).forces.append((objectUniqueId, linkIndex, tuple(float(v) for v in forceObj),
                                       tuple(float(v) for v in posObj), flags))


def applyExternalTorque(objectUniqueId, linkIndex, torqueObj, flags, physicsClientId=0):
    _c(physicsClientId).torques.append((objectUniqueId, linkIndex, tuple(float(v) for v in torqueObj), flags))


def pop_recorded(physicsClientId=0):
    """Stand-in extension: return and clear the recorded external forces/torques."""
    c = _c(physicsClientId)
    f, t = c.forces, c.torques
    c.forces, c.torques = [], []
    return f, t
