"""
Training / evaluation input preparation on the host (reference: helper/loader.py).

Only what the hot path's callers need: `build_input_image` / `build_image_set` (used by
DCSCN.do_for_evaluate, loader.py:23-67), an in-RAM grid-patch data set (`BatchDataSets`,
loader.py:70-275 without the on-disk BMP cache) and the random-crop data set (`DynamicDataSets`,
loader.py:278-355).  Data loading is CPU work outside the replaced path.
"""

import logging
import random

import numpy as np

from helper import utilty as util


def build_image_set(file_path, channels=1, scale=1, convert_ycbcr=True, resampling_method="bicubic",
                    print_console=True):
    """loader.py:23-33 -> (input LR, bicubic-upscaled LR, true HR)."""
    true_image = util.set_image_alignment(util.load_image(file_path, print_console=print_console), scale)
    if channels == 1 and true_image.shape[2] == 3 and convert_ycbcr:
        true_image = util.convert_rgb_to_y(true_image)
    input_image = util.resize_image_by_pil(true_image, 1.0 / scale, resampling_method=resampling_method)
    input_interpolated_image = util.resize_image_by_pil(input_image, scale, resampling_method=resampling_method)
    return input_image, input_interpolated_image, true_image


def build_input_image(image, width=0, height=0, channels=1, scale=1, alignment=0, convert_ycbcr=True):
    """loader.py:42-67: centre-crop, align, RGB -> Y (or YCbCr), bicubic down-scale by `scale`."""
    if width != 0 and height != 0:
        if image.shape[0] != height or image.shape[1] != width:
            x = (image.shape[1] - width) // 2
            y = (image.shape[0] - height) // 2
            image = image[y: y + height, x: x + width, :]
    if alignment > 1:
        image = util.set_image_alignment(image, alignment)
    if channels == 1 and image.shape[2] == 3:
        if convert_ycbcr:
            image = util.convert_rgb_to_y(image)
    else:
        if convert_ycbcr:
            image = util.convert_rgb_to_ycbcr(image)
    if scale != 1:
        image = util.resize_image_by_pil(image, 1.0 / scale)
    return image


def load_input_image(filename, width=0, height=0, channels=1, scale=1, alignment=0, convert_ycbcr=True,
                     print_console=True):
    image = util.load_image(filename, print_console=print_console)
    return build_input_image(image, width, height, channels, scale, alignment, convert_ycbcr)


class BatchDataSets:
    """Grid patches of every image of a directory, held in RAM as uint8 like the reference
    (loader.py:236-249) and served in a shuffled order (loader.py:259-275)."""

    def __init__(self, scale, batch_dir, batch_image_size, stride_size=0, channels=1, resampling_method="bicubic"):
        self.scale = scale
        self.batch_image_size = batch_image_size
        self.stride = batch_image_size // 2 if stride_size == 0 else stride_size
        self.channels = channels
        self.resampling_method = resampling_method
        self.count = 0
        self.batch_dir = batch_dir
        self.batch_index = None
        self.input_images = self.input_interpolated_images = self.true_images = None
        self.index = 0

    def is_batch_exist(self):
        return self.input_images is not None

    def build_batch(self, data_dir):
        print("Building batch images for %s..." % self.batch_dir)
        inputs, interps, trues = [], [], []
        out_size = self.batch_image_size * self.scale
        out_stride = self.stride * self.scale
        for filename in util.get_files_in_directory(data_dir):
            input_image, interp_image, true_image = build_image_set(
                filename, channels=self.channels, resampling_method=self.resampling_method, scale=self.scale,
                print_console=False)
            a = util.get_split_images(input_image, self.batch_image_size, stride=self.stride)
            b = util.get_split_images(interp_image, out_size, stride=out_stride)
            if a is None or b is None:
                continue
            c = util.get_split_images(true_image, out_size, stride=out_stride)
            inputs.append(a)
            interps.append(b)
            trues.append(c)
        if not inputs:
            self.count = 0
            return
        # stored as uint8 files in the reference (BMP patches): same truncation here
        self.input_images = np.concatenate(inputs).astype(np.uint8)
        self.input_interpolated_images = np.concatenate(interps).astype(np.uint8)
        self.true_images = np.concatenate(trues).astype(np.uint8)
        self.count = self.input_images.shape[0]
        print("%d mini-batch images are built(saved)." % self.count)

    def load_batch_counts(self):
        pass

    def load_all_batch_images(self):
        print("Allocating memory for all batch images.")

    def init_batch_index(self, shuffle=True):
        self.batch_index = random.sample(range(0, self.count), self.count) if shuffle else list(range(self.count))
        self.index = 0

    def get_next_image_no(self):
        if self.index >= self.count:
            self.init_batch_index()
        image_no = self.batch_index[self.index]
        self.index += 1
        return image_no

    def load_batch_image(self, max_value):
        number = self.get_next_image_no()
        if max_value == 255:
            return self.input_images[number], self.input_interpolated_images[number], self.true_images[number]
        f = max_value / 255.0
        return (np.multiply(self.input_images[number], f), np.multiply(self.input_interpolated_images[number], f),
                np.multiply(self.true_images[number], f))


class DynamicDataSets:
    """Random crops with a 50 % left-right flip (loader.py:278-355)."""

    def __init__(self, scale, batch_image_size, channels=1, resampling_method="bicubic"):
        self.scale = scale
        self.batch_image_size = batch_image_size
        self.channels = channels
        self.resampling_method = resampling_method
        self.filenames = []
        self.count = 0
        self.batch_index = None
        self.index = 0

    def set_data_dir(self, data_dir):
        self.filenames = util.get_files_in_directory(data_dir)
        self.count = len(self.filenames)
        if self.count <= 0:
            logging.error("Data Directory is empty.")
            exit(-1)

    def init_batch_index(self):
        self.batch_index = random.sample(range(0, self.count), self.count)
        self.index = 0

    def get_next_image_no(self):
        if self.index >= self.count:
            self.init_batch_index()
        image_no = self.batch_index[self.index]
        self.index += 1
        return image_no

    def load_batch_image(self, max_value):
        """loader.py:310-330"""
        image = None
        while image is None:
            image = self.load_random_patch(self.filenames[self.get_next_image_no()])
        if random.randrange(2) == 0:
            image = np.fliplr(image)
        input_image = util.resize_image_by_pil(image, 1 / self.scale)
        input_bicubic_image = util.resize_image_by_pil(input_image, self.scale)
        if max_value != 255:
            scale = max_value / 255.0
            input_image = np.multiply(input_image, scale)
            input_bicubic_image = np.multiply(input_bicubic_image, scale)
            image = np.multiply(image, scale)
        return input_image, input_bicubic_image, image

    def load_random_patch(self, filename):
        """loader.py:332-355"""
        image = util.load_image(filename, print_console=False)
        height, width = image.shape[0:2]
        load_batch_size = self.batch_image_size * self.scale
        if height < load_batch_size or width < load_batch_size:
            print("Error: %s should have more than %d x %d size." % (filename, load_batch_size, load_batch_size))
            return None
        if height == load_batch_size:
            y = 0
        else:
            y = random.randrange(height - load_batch_size)
        if width == load_batch_size:
            x = 0
        else:
            x = random.randrange(width - load_batch_size)
        image = image[y:y + load_batch_size, x:x + load_batch_size, :]
        image = build_input_image(image, channels=self.channels, convert_ycbcr=True)
        return image
