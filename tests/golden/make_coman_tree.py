"""Generates tests/golden/robots/coman_tree.npz: the kinematic tree table (opensot_b200.kinematics.load_urdf) of the
reference's COMAN model, /root/reference/tests/robots/coman/coman.urdf -- a data fixture (joint origins, axes,
link masses), so that the GPU tests and bench.py need no file of the reference at run time."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from opensot_b200 import kinematics as K  # noqa: E402

t = K.load_urdf("/root/reference/tests/robots/coman/coman.urdf")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "robots", "coman_tree.npz"), names=np.array(t.names), links=np.array(t.links),
                    parent=t.parent, type=t.type, dof=t.dof, axis=t.axis, origin=t.origin, mass=t.mass, com=t.com,
                    dof_names=np.array(t.dof_names), q_min=t.q_min, q_max=t.q_max, qdot_max=t.qdot_max)
print("coman_tree.npz: %d joints, %d dof, %.3f kg" % (t.n_joints, t.n_dof, t.mass.sum()))
