"""PoseBin record, drop-in for the reference class (PoseBin.py:4-85).

In the B200 flow the bins are produced in bulk by the voting kernel (csrc/sod_hough.cu) and
materialised as these objects for the caller; the mutators below keep the reference's semantics
for code that drives a bin by hand.
"""
import numpy as np


class PoseBin:

    def __init__(self, pose, img_size=(100, 100), votes=0, keypoint_pairs=[], mean=(0, 0, 0, 0)):
        # the mutable default is part of the reference's signature (PoseBin.py:6)
        self.pose = pose                    # (i_x, i_y, i_theta, i_sigma)
        self.img_size = img_size            # (width, height), running mean
        self.votes = votes
        self.keypoint_pairs = keypoint_pairs
        self.centroid = mean[0], mean[1]
        self.angle = mean[2]
        self.scale = mean[3]
        self.affine_parameters = []

    # ---- running means: (old * votes + new) / (votes + 1), PoseBin.py:19-43
    def _blend(self, old, new):
        return (old * self.votes + new) / (self.votes + 1)

    def update_centroid(self, center):
        self.centroid = (self._blend(self.centroid[0], center[0]), self._blend(self.centroid[1], center[1]))

    def update_angle(self, alpha):
        self.angle = self._blend(self.angle, alpha)

    def update_scale(self, scale):
        self.scale = self._blend(self.scale, scale)

    def update_img_size(self, img_size):
        self.img_size = (self._blend(self.img_size[0], img_size[0]), self._blend(self.img_size[1], img_size[1]))

    def add_keypoint_pair(self, pair):
        self.keypoint_pairs.append(pair)

    def add_vote(self, new_votes=1):
        self.votes += new_votes

    def update_posebin(self, object_pose, img_size, keypoint_pair):
        """One more vote (PoseBin.py:45-51): means first (they use the old count), then the pair."""
        self.update_centroid((object_pose[0], object_pose[1]))
        self.update_angle(object_pose[2])
        self.update_scale(object_pose[3])
        self.update_img_size(img_size)
        self.add_keypoint_pair(keypoint_pair)
        self.add_vote()

    def get_pts(self):
        """(query_x, query_y, model_x, model_y) arrays of the bin's pairs (PoseBin.py:56-67)."""
        q = np.array([[p[1].pt[0], p[1].pt[1]] for p in self.keypoint_pairs], dtype=float).reshape(-1, 2)
        m = np.array([[p[0].pt[0], p[0].pt[1]] for p in self.keypoint_pairs], dtype=float).reshape(-1, 2)
        return q[:, 0], q[:, 1], m[:, 0], m[:, 1]

    def remove_keypoint_pair(self, pair):
        self.keypoint_pairs.remove(pair)

    def is_same_pose(self, pose):
        return pose == self.pose

    def __eq__(self, other):
        return isinstance(other, PoseBin) and other.pose == self.pose

    __hash__ = None  # as in the reference: defining __eq__ alone makes instances unhashable

    def __repr__(self):
        return "[" + str(self.pose) + ", " + str(self.votes) + ", " + str(self.img_size) + "]"

    def __str__(self):
        return "Pose: " + str(self.pose)
