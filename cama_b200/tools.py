"""Small host utilities, call-compatible with /root/reference/cama/tools.py.

``VideoGenerator`` pipes the 2x3 camera mosaic to an ffmpeg child process exactly like the
reference; the ``ffmpeg`` (ffmpeg-python) module is imported lazily so that everything else in
the package works where it is not installed.
"""
from __future__ import annotations

import json

import numpy as np

MOSAIC_ROWS = (("camera_front_left", "camera_front", "camera_front_right"),
               ("camera_rear_left", "camera_rear", "camera_rear_right"))


def load_json(filename):
    with open(filename, "r") as fh:
        return json.load(fh)


def concate_image(image_dict):
    """{camera: HxWx3} -> 2H x 3W x 3 mosaic, front cameras on top (reference tools.py:22-25)."""
    rows = [np.concatenate([image_dict[name] for name in row], axis=1) for row in MOSAIC_ROWS]
    return np.concatenate(rows, axis=0)


class VideoGenerator:
    def __init__(self, output_video_path, output_shape=(2880, 1080)):
        import ffmpeg       # ffmpeg-python; only needed when a video is actually written
        self.writer = (
            ffmpeg.input('pipe:', format='rawvideo', pix_fmt='bgr24', s=f'{output_shape[0]}x{output_shape[1]}')
            .output(output_video_path, pix_fmt='yuv420p', vcodec='libx264', r=10, loglevel='quiet')
            .overwrite_output()
            .run_async(pipe_stdin=True)
        )

    def concate_image(self, image_dict):
        return concate_image(image_dict)

    def add_frame(self, image):
        self.writer.stdin.write(image.astype(np.uint8).tobytes())

    def add_frame_from_dict(self, image_dict):
        self.add_frame(self.concate_image(image_dict))

    def __del__(self):
        writer = getattr(self, "writer", None)
        if writer is not None:
            writer.stdin.close()
            writer.wait()
