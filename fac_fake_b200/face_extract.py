"""``FaceExtractorEngine`` — host-side mirror of the reference's ``FaceExtractor`` with the tiling,
the detector, the untiling / NMS / margin and the crop resize on the GPU.

Used exactly where the reference uses its class
(/root/reference/CViT-main/helpers/helpers_face_extract_1.py:7-124, cvit_prediction.py:124-149):

    video_reader = VideoReader()
    face_extractor = FaceExtractorEngine(lambda p: video_reader.read_random_frames(p, num_frames=15), facedet)
    faces = face_extractor.process_video(video_path)            # same list of per-frame dictionaries as the reference

``facedet`` is a ``BlazeFaceEngine``.  What runs where:

=============================================  ==========================================================
reference (helpers_face_extract_1.py)          here
=============================================  ==========================================================
``video_read_fn`` (cv2 decode)                 the caller's function, unchanged (no NVDEC in this image)
``_tile_frames`` :139-205                      ``ff_blazeface_tile_frames`` (INTER_AREA to 128x128, bit-exact)
``facedet.predict_on_batch`` :87               ``ff_blazeface_predict`` (network + decode)
``_resize_detections`` / ``_untile_detections``
/ ``facedet.nms`` / ``_add_margin_to_detections``
/ ``_crop_faces`` rectangle :207-312           ``ff_blazeface_frame_faces`` (one warp per frame)
crop + ``cv2.resize(224)`` + ``cvtColor``
(cvit_prediction.py:141-142)                   ``extract_crops_device``: views of the device frame -> ``ff_preprocess_crops``
=============================================  ==========================================================

Between the upload of the decoded frames and the CViT scores only the per-frame detections ([F,16,17] floats, the crop
rectangles and the counts) travel back to the host — to build the reference's result dictionaries and the per-video
offsets ``ff_cvit_predict`` needs.  Frames whose candidate / face count exceeds the kernel's limits (64 / 16) are
handled by the reference algorithm on the host.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .blazeface import BlazeFaceEngine
from .engine import CViTEngine, _stream_ptr


class FaceExtractorEngine:
    """Drop-in for ``FaceExtractor(video_read_fn, facedet)`` (helpers_face_extract_1.py:7-21)."""

    margin = 0.2                                   # helpers_face_extract_1.py:110

    def __init__(self, video_read_fn: Callable, facedet: BlazeFaceEngine):
        self.video_read_fn = video_read_fn
        self.facedet = facedet

    # ------------------------------------------------------------------ device pipeline for one video's frames
    def _tiles_per_frame(self, H: int, W: int) -> int:
        return 3 if W > H else 1                   # :185-189

    def detect_frames_device(self, frames: torch.Tensor):
        """frames: DEVICE uint8 [F,H,W,3] -> (faces [F,16,17] fp32, boxes [F,16,4] int32, counts [F] int32), all on the
        device, in frame coordinates (what the reference holds after ``self.facedet.nms(detections)``, :101-104, plus the
        margin-expanded integer crop rectangles of :110-111)."""
        fd = self.facedet
        fd._ensure_ready()
        dev = fd._device
        F, H, W, _ = frames.shape
        T = self._tiles_per_frame(H, W)
        tiles = torch.empty((F * T, 128, 128, 3), dtype=torch.uint8, device=dev)
        faces = torch.empty((F, 16, 17), dtype=torch.float32, device=dev)
        boxes = torch.empty((F, 16, 4), dtype=torch.int32, device=dev)
        counts = torch.empty((F,), dtype=torch.int32, device=dev)
        if F == 0:
            return faces, boxes, counts, tiles
        with torch.cuda.device(dev):
            st = C.c_void_p(_stream_ptr(dev))
            fd._check(fd._lib.ff_blazeface_tile_frames(fd._h, C.c_void_p(frames.data_ptr()), F, H, W, C.c_void_p(tiles.data_ptr()), st),
                      "ff_blazeface_tile_frames")
            dense = torch.cat([fd.predict_dense(tiles[i:i + fd._max_tiles]) for i in range(0, F * T, fd._max_tiles)])
            fd._check(fd._lib.ff_blazeface_frame_faces(fd._h, C.c_void_p(dense.data_ptr()), F, H, W, C.c_float(fd.min_score_thresh),
                                                       C.c_float(fd.min_suppression_threshold), C.c_float(self.margin),
                                                       C.c_void_p(faces.data_ptr()), C.c_void_p(boxes.data_ptr()),
                                                       C.c_void_p(counts.data_ptr()), st), "ff_blazeface_frame_faces")
        self._last_dense = dense
        return faces, boxes, counts, tiles

    # the reference algorithm on the host, for frames the kernel declined (count == -1)
    def _frame_on_host(self, dense_frame: torch.Tensor, H: int, W: int) -> Tuple[torch.Tensor, np.ndarray]:
        fd = self.facedet
        split = min(H, W)
        x_step = (W - split) // 2
        dets = []
        for t in range(dense_frame.shape[0]):
            d = dense_frame[t]
            d = d[d[:, 16] >= fd.min_score_thresh].clone()
            for k in range(2):                                      # _resize_detections :222-224
                d[:, k * 2] = (d[:, k * 2] * 128 - 0) * (split / 128)
                d[:, k * 2 + 1] = (d[:, k * 2 + 1] * 128 - 0) * (split / 128)
            for k in range(2, 8):                                   # :227-229
                d[:, k * 2] = (d[:, k * 2] * 128 - 0) * (split / 128)
                d[:, k * 2 + 1] = (d[:, k * 2 + 1] * 128 - 0) * (split / 128)
            x = t * x_step                                          # _untile_detections :255-262 (y stays 0)
            if d.shape[0] > 0:
                for k in range(2):
                    d[:, k * 2 + 1] += x
                for k in range(2, 8):
                    d[:, k * 2] += x
            dets.append(d)
        merged = fd._weighted_non_max_suppression(torch.cat(dets))
        det = torch.stack(merged) if merged else torch.zeros((0, 17))
        return det, self._margin_boxes(det, W, H)

    def _margin_boxes(self, det: torch.Tensor, W: int, H: int) -> np.ndarray:
        """_add_margin_to_detections :274-294 + the int truncation of _crop_faces :307."""
        offset = torch.round(self.margin * (det[:, 2] - det[:, 0]))
        ymin = torch.clamp(det[:, 0] - offset * 2, min=0)
        xmin = torch.clamp(det[:, 1] - offset, min=0)
        ymax = torch.clamp(det[:, 2] + offset, max=H)
        xmax = torch.clamp(det[:, 3] + offset, max=W)
        return torch.stack([ymin, xmin, ymax, xmax], 1).numpy().astype(int).reshape(-1, 4)

    def _frames_to_lists(self, frames_dev: torch.Tensor):
        """Per frame: (detections [k,17] CPU tensor, crop rectangles int [k,4])."""
        F, H, W, _ = frames_dev.shape
        faces, boxes, counts, _ = self.detect_frames_device(frames_dev)
        faces_h, boxes_h, counts_h = faces.cpu(), boxes.cpu().numpy(), counts.cpu().tolist()
        T = self._tiles_per_frame(H, W)
        out = []
        for f, k in enumerate(counts_h):
            if k >= 0:
                out.append((faces_h[f, :k].clone(), boxes_h[f, :k].astype(int)))
            else:
                out.append(self._frame_on_host(self._last_dense[f * T:(f + 1) * T].cpu(), H, W))
        return out

    # ------------------------------------------------------------------ the reference's public methods
    def process_videos(self, input_dir, filenames, video_idxs):
        """helpers_face_extract_1.py:23-118: same list of per-frame dictionaries (video_idx, frame_idx, frame_w, frame_h,
        faces = NumPy crops of the ORIGINAL frame, scores)."""
        dev = self.facedet._device or torch.device("cuda", torch.cuda.current_device())
        result = []
        for video_idx in video_idxs:
            video_path = os.path.join(input_dir, filenames[video_idx])
            got = self.video_read_fn(video_path)
            if got is None:
                continue
            my_frames, my_idxs = got
            frames_dev = torch.from_numpy(np.ascontiguousarray(my_frames)).to(dev)
            F, H, W, _ = my_frames.shape
            for i, (det, rect) in enumerate(self._frames_to_lists(frames_dev)):
                faces = [my_frames[i][y0:y1, x0:x1, :] for (y0, x0, y1, x1) in rect]      # _crop_faces :305-311
                result.append({"video_idx": video_idx, "frame_idx": my_idxs[i], "frame_w": W, "frame_h": H,
                               "faces": faces, "scores": list(det[:, 16].numpy())})
        return result

    def process_video(self, video_path):
        """helpers_face_extract_1.py:120-124."""
        return self.process_videos(os.path.dirname(video_path), [os.path.basename(video_path)], [0])

    def remove_large_crops(self, crops, pct=0.1):
        """helpers_face_extract_1.py:318-344 (the reference compares against 0.1, not ``pct``: kept)."""
        for frame_data in crops:
            video_area = frame_data["frame_w"] * frame_data["frame_h"]
            keep = [j for j, face in enumerate(frame_data["faces"]) if face.shape[0] * face.shape[1] / video_area < 0.1]
            frame_data["faces"] = [frame_data["faces"][j] for j in keep]
            frame_data["scores"] = [frame_data["scores"][j] for j in keep]

    def keep_only_best_face(self, crops):
        """helpers_face_extract_1.py:346-359."""
        for frame_data in crops:
            if len(frame_data["faces"]) > 0:
                frame_data["faces"] = frame_data["faces"][:1]
                frame_data["scores"] = frame_data["scores"][:1]

    # ------------------------------------------------------------------ fused device path: frames -> 224x224 crops on the GPU
    def extract_crops_device(self, frames: torch.Tensor, model: CViTEngine, max_crops: Optional[int] = None) -> torch.Tensor:
        """What ``face_blaze`` builds on the host (cvit_prediction.py:124-149) without the pixels leaving the GPU:
        frames DEVICE uint8 [F,H,W,3] (RGB as decoded by VideoReader) -> DEVICE uint8 [n,224,224,3] crops
        (``cv2.resize(INTER_AREA)`` + ``cvtColor(RGB2BGR)`` of every non-empty face rectangle, frame by frame)."""
        views: List[torch.Tensor] = []
        for f, (_, rect) in enumerate(self._frames_to_lists(frames)):
            for (y0, x0, y1, x1) in rect:
                if y1 > y0 and x1 > x0 and (max_crops is None or len(views) < max_crops):       # `face.size > 0`
                    views.append(frames[f, y0:y1, x0:x1, :])
        if not views:
            return torch.empty((0, 224, 224, 3), dtype=torch.uint8, device=frames.device)
        return model.preprocess_crops(views, swap_rb=True)
