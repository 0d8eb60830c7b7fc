"""
Host-side utilities of the DCSCN drop-in (image I/O, colour conversion, PIL bicubic, PSNR / SSIM,
ensemble flips, logging).  Behaviour follows the reference's helper/utilty.py function by function
(cited below) so PSNR numbers are apples-to-apples; TensorFlow / imageio / scikit-image are not needed.
The module keeps the reference's (misspelled) name so `from helper import utilty as util` still works.
"""

import datetime
import logging
import math
import os

import numpy as np
from PIL import Image


class LoadError(Exception):
    def __init__(self, message):
        self.message = message


# ---------------------------------------------------------------- files ----

def make_dir(directory):
    """utilty.py:56-58"""
    os.makedirs(directory, exist_ok=True)


def clean_dir(path):
    """utilty.py:81-94: remove everything inside `path`, keep `path`."""
    if not os.path.isdir(path):
        return
    for entry in os.listdir(path):
        full = os.path.join(path, entry)
        try:
            if os.path.isdir(full):
                clean_dir(full)
                os.rmdir(full)
            else:
                os.remove(full)
        except OSError as error:
            print("OS error: {0}".format(error))


def delete_dir(directory):
    """utilty.py:61-64"""
    if os.path.exists(directory):
        clean_dir(directory)
        os.rmdir(directory)


def get_files_in_directory(path):
    """utilty.py:67-71 (sorted here, the reference takes os.listdir order; averages do not depend on it)."""
    if not path.endswith('/'):
        path = path + "/"
    return [path + f for f in sorted(os.listdir(path))
            if os.path.isfile(os.path.join(path, f)) and not f.startswith('.')]


def set_logging(filename, stream_log_level, file_log_level, tf_log_level=None):
    """utilty.py:97-110 (the TensorFlow verbosity argument is accepted and ignored)."""
    stream_log = logging.StreamHandler()
    stream_log.setLevel(stream_log_level)
    file_log = logging.FileHandler(filename=filename)
    file_log.setLevel(file_log_level)
    logger = logging.getLogger()
    logger.handlers = []
    logger.addHandler(stream_log)
    logger.addHandler(file_log)
    logger.setLevel(min(stream_log_level, file_log_level))


def get_now_date():
    """utilty.py:475-477"""
    d = datetime.datetime.today()
    return "%s/%s/%s %s:%s:%s" % (d.year, d.month, d.day, d.hour, d.minute, d.second)


# ---------------------------------------------------------------- image I/O ----

def save_image(filename, image, print_console=True):
    """utilty.py:113-131: astype(uint8) truncation without clipping, like the reference."""
    if len(image.shape) >= 3 and image.shape[2] == 1:
        image = image.reshape(image.shape[0], image.shape[1])
    directory = os.path.dirname(filename)
    if directory != "":
        os.makedirs(directory, exist_ok=True)
    data = image.astype(np.uint8)
    if data.ndim == 3 and data.shape[2] == 3:
        Image.fromarray(data, mode="RGB").save(filename)
    else:
        Image.fromarray(data).save(filename)
    if print_console:
        print("Saved [%s]" % filename)


def load_image(filename, width=0, height=0, channels=0, alignment=0, print_console=True):
    """utilty.py:242-266 (imageio.imread -> PIL: identical decoded pixels for the PNG/BMP/JPEG inputs)."""
    if not os.path.isfile(filename):
        raise LoadError("File not found [%s]" % filename)
    im = Image.open(filename)
    if im.mode == "P":
        im = im.convert("RGB")
    elif im.mode not in ("L", "RGB", "RGBA"):
        im = im.convert("RGB")
    image = np.atleast_3d(np.asarray(im))
    if (width != 0 and image.shape[1] != width) or (height != 0 and image.shape[0] != height):
        raise LoadError("Attributes mismatch")
    if channels != 0 and image.shape[2] != channels:
        raise LoadError("Attributes mismatch")
    if alignment != 0 and ((width % alignment) != 0 or (height % alignment) != 0):
        raise LoadError("Attributes mismatch")
    if image.shape[2] >= 4:
        image = image[:, :, 0:3]
    if print_console:
        print("Loaded [%s]: %d x %d x %d" % (filename, image.shape[1], image.shape[0], image.shape[2]))
    return image


# ---------------------------------------------------------------- colour ----

_Y_ROW = np.array([[65.738 / 256.0, 129.057 / 256.0, 25.064 / 256.0]])
_YCBCR = np.array([[65.738 / 256.0, 129.057 / 256.0, 25.064 / 256.0],
                   [-37.945 / 256.0, -74.494 / 256.0, 112.439 / 256.0],
                   [112.439 / 256.0, -94.154 / 256.0, -18.285 / 256.0]])
_RGB = np.array([[298.082 / 256.0, 0, 408.583 / 256.0],
                 [298.082 / 256.0, -100.291 / 256.0, -208.120 / 256.0],
                 [298.082 / 256.0, 516.412 / 256.0, 0]])


def convert_rgb_to_y(image):
    """utilty.py:142-149"""
    if len(image.shape) <= 2 or image.shape[2] == 1:
        return image
    return image.dot(_Y_ROW.T) + 16.0


def convert_rgb_to_ycbcr(image):
    """utilty.py:152-165"""
    if len(image.shape) < 2 or image.shape[2] == 1:
        return image
    ycbcr = image.dot(_YCBCR.T)
    ycbcr[:, :, 0] += 16.0
    ycbcr[:, :, [1, 2]] += 128.0
    return ycbcr


def convert_ycbcr_to_rgb(ycbcr_image):
    """utilty.py:168-179"""
    shifted = np.zeros([ycbcr_image.shape[0], ycbcr_image.shape[1], 3])
    shifted[:, :, 0] = ycbcr_image[:, :, 0] - 16.0
    shifted[:, :, [1, 2]] = ycbcr_image[:, :, [1, 2]] - 128.0
    return shifted.dot(_RGB.T)


def convert_y_and_cbcr_to_rgb(y_image, cbcr_image):
    """utilty.py:182-193"""
    if len(y_image.shape) <= 2:
        y_image = y_image.reshape(y_image.shape[0], y_image.shape[1], 1)
    if len(y_image.shape) == 3 and y_image.shape[2] == 3:
        y_image = y_image[:, :, 0:1]
    ycbcr = np.zeros([y_image.shape[0], y_image.shape[1], 3])
    ycbcr[:, :, 0] = y_image[:, :, 0]
    ycbcr[:, :, 1:3] = cbcr_image[:, :, 0:2]
    return convert_ycbcr_to_rgb(ycbcr)


# ---------------------------------------------------------------- geometry ----

def set_image_alignment(image, alignment):
    """utilty.py:196-208: crop to a multiple of `alignment`, drop alpha."""
    alignment = int(alignment)
    height = (image.shape[0] // alignment) * alignment
    width = (image.shape[1] // alignment) * alignment
    if image.shape[1] != width or image.shape[0] != height:
        image = image[:height, :width, :]
    if len(image.shape) >= 3 and image.shape[2] >= 4:
        image = image[:, :, 0:3]
    return image


def resize_image_by_pil(image, scale, resampling_method="bicubic"):
    """utilty.py:211-239.  Float single-channel arrays go through PIL as mode 'F' (no quantisation),
    uint8 ones as mode 'L' - both exactly like the reference."""
    width, height = image.shape[1], image.shape[0]
    new_width = int(width * scale)
    new_height = int(height * scale)
    method = {"bicubic": Image.BICUBIC, "bilinear": Image.BILINEAR, "nearest": Image.NEAREST}.get(
        resampling_method, Image.LANCZOS)
    if len(image.shape) == 3 and image.shape[2] in (3, 4):
        im = Image.fromarray(image, "RGB").resize([new_width, new_height], resample=method)
        return np.asarray(im)
    im = Image.fromarray(image.reshape(height, width)).resize([new_width, new_height], resample=method)
    return np.asarray(im).reshape(new_height, new_width, 1)


def flip(image, flip_type, invert=False):
    """utilty.py:595-617: the 8 self-ensemble transforms and their inverses."""
    if flip_type == 0:
        return image
    if flip_type == 1:
        return np.flipud(image)
    if flip_type == 2:
        return np.fliplr(image)
    if flip_type == 3:
        return np.flipud(np.fliplr(image))
    if flip_type == 4:
        return np.rot90(image, 1 if invert is False else -1)
    if flip_type == 5:
        return np.rot90(image, -1 if invert is False else 1)
    if flip_type == 6:
        return np.flipud(np.rot90(image)) if invert is False else np.rot90(np.flipud(image), -1)
    if flip_type == 7:
        return np.flipud(np.rot90(image, -1)) if invert is False else np.rot90(np.flipud(image), 1)
    raise ValueError("flip_type must be 0..7")


def get_split_images(image, window_size, stride=None, enable_duplicate=False):
    """utilty.py:286-327: sliding windows [count, window, window, 1] (grid patches for build_batch)."""
    if len(image.shape) == 3 and image.shape[2] == 1:
        image = image.reshape(image.shape[0], image.shape[1])
    window_size = int(window_size)
    stride = window_size if stride is None else int(stride)
    height, width = image.shape
    if height < window_size or width < window_size:
        return None
    ys = range(0, height - window_size + 1, stride)
    xs = range(0, width - window_size + 1, stride)
    windows = [image[y:y + window_size, x:x + window_size] for y in ys for x in xs]
    if enable_duplicate:
        if (height - window_size) % stride != 0:
            for x in range(0, width - window_size, stride):
                windows.append(image[height - window_size - 1:height - 1, x:x + window_size])
        if (width - window_size) % stride != 0:
            for y in range(0, height - window_size, stride):
                windows.append(image[y:y + window_size, width - window_size - 1:width - 1])
    out = np.stack(windows, axis=0)
    return out.reshape(out.shape[0], window_size, window_size, 1)


# ---------------------------------------------------------------- metrics ----

def trim_image_as_file(image):
    """utilty.py:501-506: what a saved 8-bit file would hold."""
    image = np.clip(np.rint(image), 0, 255)
    if image.dtype != np.float32:
        image = image.astype(np.float32)
    return image


def get_loss_image(image1, image2, scale=1.0, border_size=0):
    """utilty.py:480-498"""
    if len(image1.shape) == 2:
        image1 = image1.reshape(image1.shape[0], image1.shape[1], 1)
    if len(image2.shape) == 2:
        image2 = image2.reshape(image2.shape[0], image2.shape[1], 1)
    if image1.shape != image2.shape:
        return None
    loss_image = np.minimum(np.square(trim_image_as_file(image1) - trim_image_as_file(image2)) * scale, 255.0)
    if border_size > 0:
        loss_image = loss_image[border_size:-border_size, border_size:-border_size, :]
    return loss_image


def _ssim_columns(a, b):
    """skimage.metrics.structural_similarity(a, b, win_size=11, gaussian_weights=True, multichannel=True,
    K1=0.01, K2=0.03, sigma=1.5, data_range=255) on 2-D inputs, as the reference calls it (utilty.py:534-535):
    `multichannel=True` makes the LAST axis (image columns) the channel axis, so each column is scored as a
    1-D signal and the results are averaged.  Restated with scipy's gaussian_filter1d."""
    from scipy.ndimage import gaussian_filter1d
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    if a.shape[0] < 11:
        return float("nan")

    def filt(v):
        return gaussian_filter1d(v, 1.5, axis=0, truncate=3.5, mode="reflect")

    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    ux, uy = filt(a), filt(b)
    vx = filt(a * a) - ux * ux
    vy = filt(b * b) - uy * uy
    vxy = filt(a * b) - ux * uy
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    pad = 5
    return float(np.mean(s[pad:-pad, :]))


def compute_psnr_and_ssim(image1, image2, border_size=0):
    """utilty.py:509-536: round + clip to 0..255, shave `border_size`, PSNR with data_range 255
    (== 10*log10(255^2/mse), float64 like skimage's peak_signal_noise_ratio) and SSIM."""
    if len(image1.shape) == 2:
        image1 = image1.reshape(image1.shape[0], image1.shape[1], 1)
    if len(image2.shape) == 2:
        image2 = image2.reshape(image2.shape[0], image2.shape[1], 1)
    if image1.shape != image2.shape:
        return None
    image1 = trim_image_as_file(image1)
    image2 = trim_image_as_file(image2)
    if border_size > 0:
        image1 = image1[border_size:-border_size, border_size:-border_size, :]
        image2 = image2[border_size:-border_size, border_size:-border_size, :]
    if image1.shape[2] == 1:
        image1 = image1[:, :, 0]
        image2 = image2[:, :, 0]
    err = np.mean((image1.astype(np.float64) - image2.astype(np.float64)) ** 2)
    psnr = float("inf") if err == 0 else 10.0 * math.log10(255.0 * 255.0 / err)
    ssim = _ssim_columns(image1, image2) if image1.ndim == 2 else float("nan")
    return psnr, ssim


def get_psnr(mse, max_value=255.0):
    """utilty.py:561-566"""
    if mse is None or mse == float('Inf') or mse == 0:
        return 0
    return 20 * math.log(max_value / math.sqrt(mse), 10)


def print_num_of_total_parameters(model=None, output_detail=False, output_to_logging=False):
    """utilty.py:569-592; the variables come from the engine instead of tf.trainable_variables()."""
    shapes = model.trainable_shapes() if model is not None else {}
    total = 0
    parts = []
    for name, shape in shapes.items():
        count = int(np.prod(shape)) if len(shape) else 1
        total += count
        parts.append("%s:0 %d, " % (name, count) if len(shape) == 1 else "%s:0 %s=%d, " % (name, str(tuple(shape)), count))
    emit = logging.info if output_to_logging else print
    if output_detail:
        emit("".join(parts))
    emit("Total %d variables, %s params" % (len(shapes), "{:,}".format(total)))
