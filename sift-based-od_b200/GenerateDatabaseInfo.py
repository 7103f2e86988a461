"""Model-database builder, drop-in for GenerateDatabaseInfo.py:1-37.

`save_object` and the pickle row layout [temp_kp, des, img_size, centroid, path] are unchanged, so
main.Main.get_query_features reads databases written by either implementation.  Unlike the
reference, importing this module does nothing; run it as a script or call build_database().
OpenCV SIFT stays the feature extractor (input stage, out of the GPU path's scope).
"""
import os
import pickle
import sys

import cv2

from SiftHelperFunctions import get_centroid, make_temp_kp

MAX_DIM = 1500  # reference :23


def save_object(obj, filename):
    with open(filename, 'wb') as outp:  # overwrites any existing file
        pickle.dump(obj, outp, pickle.HIGHEST_PROTOCOL)


def build_database(image_dir, out_file='training_data.pkl', packed_file=None):
    """Writes the reference's pickle; with packed_file also the array form that loads straight into
    HBM (sod_b200/database.py), which main.Main.load_database reads as well."""
    sift = cv2.SIFT_create()
    data = []
    for name in os.listdir(image_dir):
        img = cv2.imread(os.path.join(image_dir, name))
        if img is None:
            continue
        img = cv2.resize(img, (MAX_DIM, int(MAX_DIM * img.shape[0] / img.shape[1])))
        img_size = (img.shape[1], img.shape[0])
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        kp, des = sift.detectAndCompute(gray, None)
        data.append([make_temp_kp(kp), des, img_size, get_centroid(kp), os.path.join(image_dir, name)])
    save_object(data, out_file)
    if packed_file is not None:
        from sod_b200.database import PackedDatabase
        PackedDatabase.from_reference_rows(data).save(packed_file)
    return data


if __name__ == "__main__":
    build_database(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'training_data.pkl',
                   sys.argv[3] if len(sys.argv) > 3 else None)
