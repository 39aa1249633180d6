"""Host-side image cropping of the SiamFC tracker: the data-format side of the path, kept on the CPU with OpenCV like
the reference (projects/siamfc-pytorch/siamfc/ops.py:87-104 ``crop_and_resize(faster=True)`` and
image_utils.py:7-76 ``get_cropped_input``).  The crops are tiny (<= 255 x 255 x 3 bytes each) and feed one pinned
host->device copy per frame."""
import numbers

import cv2
import numpy as np


def _cropped_input(image, bbox, out_size, interpolation, pad_color):
    """Crop ``bbox`` (x1, y1, x2, y2; may leave the image), resize the in-image part so that the full box would map to
    out_size x out_size, and pad the rest with ``pad_color`` (image_utils.py:7-76 with padScale = 1)."""
    bbox = np.array(bbox)
    width = float(bbox[2] - bbox[0])
    height = float(bbox[3] - bbox[1])
    im_shape = np.array(image.shape)
    if len(im_shape) < 3:
        image = image[:, :, np.newaxis]
    xc = float(bbox[0] + bbox[2]) / 2
    yc = float(bbox[1] + bbox[3]) / 2
    box_on = np.array([xc - width / 2, yc - height / 2, xc + width / 2, yc + height / 2], dtype=np.float64)
    box_on = np.round(box_on).astype(int)
    box_wh = np.array([box_on[2] - box_on[0], box_on[3] - box_on[1]])
    patch = image[max(box_on[1], 0):min(box_on[3], im_shape[0]), max(box_on[0], 0):min(box_on[2], im_shape[1]), :]
    bounded = np.clip(box_on, 0, im_shape[[1, 0, 1, 0]])
    bounded_wh = np.array([bounded[2] - bounded[0], bounded[3] - bounded[1]])
    if patch.shape[0] == 0 or patch.shape[1] == 0:
        return np.zeros((int(out_size), int(out_size), 3), dtype=patch.dtype)
    patch = cv2.resize(patch, (max(1, int(np.round(out_size * bounded_wh[0] / box_wh[0]))),
                               max(1, int(np.round(out_size * bounded_wh[1] / box_wh[1])))),
                       interpolation=interpolation)
    if len(patch.shape) < 3:
        patch = patch[:, :, np.newaxis]
    patch_shape = np.array(patch.shape)
    pad = np.zeros(4, dtype=int)
    pad[:2] = np.maximum(0, -box_on[:2] * out_size / box_wh)
    pad[2:] = out_size - (pad[:2] + patch_shape[[1, 0]])
    if np.any(pad != 0):
        if len(pad[pad < 0]) > 0:
            return np.zeros((int(out_size), int(out_size), 3))
        if isinstance(pad_color, numbers.Number):
            return np.pad(patch, ((pad[1], pad[3]), (pad[0], pad[2]), (0, 0)), 'constant', constant_values=pad_color)
        return cv2.copyMakeBorder(patch, pad[1], pad[3], pad[0], pad[2], cv2.BORDER_CONSTANT, value=pad_color)
    return patch


def crop_and_resize(img, center, size, out_size, border_value=(0, 0, 0), interp=cv2.INTER_LINEAR):
    """Square crop of side ``size`` centred on ``center`` = (y, x), resized to ``out_size`` (ops.py:87-104).  Like the
    reference's fast path, the padding colour is the mean colour of ``img`` (``border_value`` is accepted and
    unused there too)."""
    size = max(2, size)
    cx, cy = float(center[1]), float(center[0])
    xyxy = [np.float32(cx) - np.float32(size) / 2.0, np.float32(cy) - np.float32(size) / 2.0,
            np.float32(cx) + np.float32(size) / 2.0, np.float32(cy) + np.float32(size) / 2.0]
    avg_color = np.mean(img, axis=(0, 1), dtype=float)
    return _cropped_input(img, xyxy, out_size, interp, avg_color)
