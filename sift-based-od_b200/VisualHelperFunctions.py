"""Display helpers, drop-in for VisualHelperFunctions.py:6-37.  matplotlib is optional: without it
the functions print the final pose and return the axes untouched."""
import math

import cv2

try:  # pragma: no cover - depends on the environment
    from matplotlib import patches
except Exception:  # matplotlib absent
    patches = None


def show_keypoints(rgb_query, keypoint_pairs, ax):
    if ax is None or patches is None:
        return ax
    query_kp = [p[1] for p in keypoint_pairs]
    ax.imshow(cv2.drawKeypoints(rgb_query, query_kp, None, flags=cv2.DRAW_MATCHES_FLAGS_DRAW_RICH_KEYPOINTS))
    return ax


def show_object(final_pose, ax):
    for (cx, cy), angle, scale, (w, h) in final_pose:
        print("final pose: centroid", (cx, cy), "angle", angle, "scale", scale)
        if ax is None or patches is None:
            continue
        sw, sh = w * scale, h * scale
        x0 = cx - (math.cos(angle) * sw / 2 - math.sin(angle) * sh / 2)
        y0 = cy - (math.sin(angle) * sw / 2 + math.cos(angle) * sh / 2)
        ax.add_patch(patches.Rectangle((x0, y0), sw, sh, angle=math.degrees(angle), linewidth=2,
                                       edgecolor="r", facecolor="none"))
        ax.plot(cx, cy, "r+")
    return ax
