"""Pipeline driver, drop-in for the reference's main.py (class Main, main.py:17-176).

Same attributes, methods and call order; the three hot methods run on the B200 through
libsod_b200.so (no CPU path):
    run_matcher               -> tcgen05 2-NN matcher + exact ratio test   (csrc/sod_match.cu)
    apply_hough_transform     -> pose / bin / vote kernels                  (csrc/sod_hough.cu)
    apply_affine_parameters   -> per-bin fit / prune fixed point            (csrc/sod_affine.cu)
OpenCV SIFT remains the feature extractor.  matplotlib is optional.
"""
import cv2
import numpy as np

from SiftHelperFunctions import *  # noqa: F401,F403
from PoseBin import *  # noqa: F401,F403
from HoughTransformHelperFunctions import *  # noqa: F401,F403
from AffineParameters import *  # noqa: F401,F403
from PostProcessing import *  # noqa: F401,F403
from VisualHelperFunctions import *  # noqa: F401,F403

from PoseBin import PoseBin
from PostProcessing import find_max_orientation, get_final_pose, group_orientation, group_position
from VisualHelperFunctions import show_keypoints, show_object
from sod_b200 import database as _database
from sod_b200 import dropin as _dropin
from sod_b200 import engine as _engine

try:  # pragma: no cover - depends on the environment
    from matplotlib import pyplot as plt
except Exception:
    plt = None

sift = None  # created on first use (the reference defines it only when run as a script)

TRAINING_DATA_PATH = '../Data_Set/Train_DataSet/DatabaseInfo/standing/training_data.pkl'


class Main:

    def __init__(self):
        self.kp = []
        self.des = []
        self.kp_query = []
        self.des_query = []
        self.rgb_query = []
        self.gray_query = []
        self.img_size_list = []
        self.img_centroid_list = []
        self.image_query_size = (0, 0)
        self.matching_keypoints = []  # (model keypoint, query keypoint, model image size, model centroid)
        self.hough_transform = {}     # pose -> PoseBin
        self.valid_bins = []
        self.keypoint_pairs = []
        self.final_pose = []
        self.ax = plt.subplots()[1] if plt is not None else None
        self._db_cache = None         # (id(des), Matcher) so repeated queries reuse the resident DB

    def get_query_features(self, path, training_data_path=TRAINING_DATA_PATH):
        global sift
        if sift is None:
            sift = cv2.SIFT_create()
        image_query = cv2.imread(path)
        self.rgb_query = cv2.cvtColor(image_query, cv2.COLOR_BGR2RGB)
        self.gray_query = cv2.cvtColor(image_query, cv2.COLOR_BGR2GRAY)
        self.kp_query, self.des_query = sift.detectAndCompute(self.gray_query, None)
        self.image_query_size = (len(self.gray_query[0]), len(self.gray_query))
        self.load_database(training_data_path)

    def load_database(self, training_data_path=TRAINING_DATA_PATH):
        """Model database: the reference pickle (main.py:50-66) or the packed array file
        GenerateDatabaseInfo.build_database(..., packed_file=...) writes (sod_b200/database.py)."""
        db = _database.PackedDatabase.open(training_data_path)
        self.img_size_list, self.img_centroid_list = db.per_keypoint_lists()
        self.kp = db.keypoints()
        self.des = db.des  # u8; the reference holds the same integers as float32

    def _matcher(self, exact=True):
        """Resident database for the exact u8 path, or for the bf16 path when either side holds
        non-integer descriptors (cached per self.des object and path)."""
        key = "u8" if exact else "f32"
        if self._db_cache is None or self._db_cache[0] is not self.des:
            self._db_cache = (self.des, {})
        cache = self._db_cache[1]
        if key not in cache:
            if exact:
                cache[key] = _engine.Matcher(_engine.prepare_db(_engine.pack_descriptors(self.des)))
            else:
                des = _engine.torch.as_tensor(np.ascontiguousarray(self.des, dtype=np.float32)).cuda()
                cache[key] = _engine.FloatMatcher(_engine.prepare_db_float(des))
        return cache[key]

    def run_matcher(self):
        """knnMatch(k=2) + ratio 0.75 (main.py:68-86).  Integer-valued descriptors (what OpenCV SIFT
        produces) take the exact u8 path; anything else the bf16 path with its stated tolerance."""
        n_train = len(self.des)
        if n_train < 2:  # the reference's `for m, n in matches` cannot unpack
            raise ValueError("not enough values to unpack (expected 2, got %d)" % n_train)
        try:
            q = _engine.pack_descriptors(self.des_query)
            idx, _, _, ok = _engine.knn_match_ratio(q, self._matcher(exact=True))
        except _engine.NonIntegerDescriptors:
            q = _engine.torch.as_tensor(np.ascontiguousarray(self.des_query, dtype=np.float32)).cuda()
            idx, _, _, ok = _engine.knn_match_ratio_float(q, self._matcher(exact=False))
        idx = idx.cpu().numpy()
        for qi in np.nonzero(ok.cpu().numpy())[0]:
            t = int(idx[qi, 0])
            self.matching_keypoints.append((self.kp[t], self.kp_query[int(qi)], self.img_size_list[t],
                                            self.img_centroid_list[t]))

    def apply_hough_transform(self, bins=15):
        """Every match votes for 2x2x2x2 adjacent pose bins (main.py:89-119)."""
        if self.hough_transform:
            raise RuntimeError("Main is single-shot: hough_transform is already populated")
        shape = self.rgb_query.shape
        self.hough_transform.update(_dropin.hough_dict(self.matching_keypoints, int(shape[1]), int(shape[0]),
                                                       int(bins), PoseBin))

    def get_valid_bins(self, threshold=5):
        self.keypoint_pairs = []
        for pose_bin in self.hough_transform.values():
            if pose_bin.votes >= threshold:
                self.keypoint_pairs.extend(pose_bin.keypoint_pairs)
                if pose_bin not in self.valid_bins:
                    self.valid_bins.append(pose_bin)

    def update_keypoint_pairs(self):
        self.keypoint_pairs.clear()
        for posebin in self.valid_bins:
            self.keypoint_pairs.extend(posebin.keypoint_pairs)

    def apply_affine_parameters(self, threshold):
        """Fit / prune every valid bin to its fixed point, keep bins with >= threshold pairs
        (main.py:139-157)."""
        pos_factor = 32
        results = _dropin.affine_run(self.valid_bins, self.image_query_size, pos_factor * 4, pos_factor * 4,
                                     int(threshold), 0)
        remaining = []
        for pose_bin, (params, keep, votes, live) in zip(self.valid_bins, results):
            if params is not None:
                pose_bin.affine_parameters = [params[0], params[1], params[2], params[3], params[4], params[5]]
            pose_bin.keypoint_pairs = [p for p, k in zip(pose_bin.keypoint_pairs, keep) if k]
            pose_bin.votes = len(pose_bin.keypoint_pairs)
            if live:
                remaining.append(pose_bin)
        self.valid_bins = remaining
        self.update_keypoint_pairs()

    def post_process(self):
        pose_cluster = group_position(self.valid_bins)
        orientation_cluster = group_orientation(pose_cluster)
        final_orientation_list = find_max_orientation(orientation_cluster)
        self.final_pose = get_final_pose(pose_cluster, final_orientation_list)

    def plot(self):
        self.ax = show_keypoints(self.rgb_query, self.keypoint_pairs, self.ax)
        self.ax = show_object(self.final_pose, self.ax)


if __name__ == "__main__":
    import sys
    main = Main()
    main.get_query_features(sys.argv[1] if len(sys.argv) > 1 else '../Data_Set/Test_DataSet/standing/random/random_6.jpg')
    main.run_matcher()
    main.apply_hough_transform(15)
    main.get_valid_bins(5)
    main.apply_affine_parameters(4)
    main.post_process()
    main.plot()
    if plt is not None:
        plt.show()
