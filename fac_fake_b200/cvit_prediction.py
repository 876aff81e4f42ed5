"""Mirror of the reference prediction API (/root/reference/CViT-main/cvit_prediction.py) with the
``CViTEngine`` in place of the PyTorch module: *video in -> forgery probability out*.

Same function names, argument meaning, sentinels and chunking as the reference:

=====================================  ==========================================================
reference (cvit_prediction.py)         here
=====================================  ==========================================================
``model = CViT(...)`` :62-70           ``configure(model=CViTEngine(...).to(dev).load_state_dict(sd))``
``predict_on_video(files, workers)``   same (:73-83) — ThreadPoolExecutor over ``predict``
``predict(filename, mtcnn)`` :153-242  same frame sampling (every 5th frame, 10 % of the clip, <= 29 crops),
                                       same <=32 chunking [0:32],[32:64],[64:90], same 0.5 sentinels
``non_empty`` :245, ``pred_sig`` :258, ``pred_tensor`` :262, ``pre_process_prediction`` :266,
``real_or_fake`` :284                  same semantics
=====================================  ==========================================================

Face detection (layer L3, dlib / MTCNN / BlazeFace in the reference) is outside the hot path
(SURVEY.md §8f-3): the detector is injected with ``configure(face_extractor=...)`` and must
behave like the reference's ``face_face_rec(frame, _) -> (faces uint8 [k,224,224,3], k)``.
The default tries ``face_recognition`` exactly like cvit_prediction.py:106-121 and raises if it
is not installed.  The crop resize itself (cv2.resize INTER_AREA + cvtColor) runs on the GPU
through ``CViTEngine.preprocess_crops`` when the extractor returns raw boxes (see
``face_boxes_extractor``).
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from time import perf_counter
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import FF_REDUCE_REFERENCE_PROBS
from .engine import CViTEngine

mean = [0.485, 0.456, 0.406]      # cvit_prediction.py:41
std = [0.229, 0.224, 0.225]       # cvit_prediction.py:42

model: Optional[CViTEngine] = None
device = "cuda"
sample = "."                      # directory of the videos (cvit_prediction.py:49)
_face_extractor: Optional[Callable] = None
verbose = False


def configure(model_: Optional[CViTEngine] = None, face_extractor: Optional[Callable] = None,
              sample_dir: Optional[str] = None, device_: Optional[str] = None, verbose_: Optional[bool] = None):
    """Set the module-level state the reference keeps in globals (cvit_prediction.py:32-70)."""
    global model, _face_extractor, sample, device, verbose
    if model_ is not None:
        model = model_
    if face_extractor is not None:
        _face_extractor = face_extractor
    if sample_dir is not None:
        sample = sample_dir
    if device_ is not None:
        device = device_
    if verbose_ is not None:
        verbose = verbose_


def predict_on_video(dfdc_filenames: Sequence[str], num_workers: int) -> List[float]:
    """cvit_prediction.py:73-83."""
    def process_file(i):
        filename = dfdc_filenames[i]
        return predict(os.path.join(sample, filename), None)

    with ThreadPoolExecutor(max_workers=num_workers) as ex:
        predictions = ex.map(process_file, range(len(dfdc_filenames)))
    return list(predictions)


def face_face_rec(frame: np.ndarray, face_tensor_face_rec=None) -> Tuple[np.ndarray, int]:
    """cvit_prediction.py:106-121 — needs the optional `face_recognition` package (dlib)."""
    try:
        import face_recognition  # type: ignore
    except ImportError as e:  # pragma: no cover - optional dependency
        raise RuntimeError("face_recognition is not installed; pass configure(face_extractor=...)") from e
    import cv2
    face_locations = face_recognition.face_locations(frame)
    temp_face = np.zeros((5, 224, 224, 3), dtype=np.uint8)
    count = 0
    for top, right, bottom, left in face_locations:
        if count < 5:
            face_image = frame[top:bottom, left:right]
            face_image = cv2.resize(face_image, (224, 224), interpolation=cv2.INTER_AREA)
            temp_face[count] = cv2.cvtColor(face_image, cv2.COLOR_RGB2BGR)
            count += 1
    if count == 0:
        return [], 0
    return temp_face[:count], count


def face_boxes_extractor(box_fn: Callable[[np.ndarray], Sequence[Tuple[int, int, int, int]]]):
    """Build an extractor from a box detector ``box_fn(frame) -> [(top, right, bottom, left), ...]``.
    The crop resize + colour swap (cvit_prediction.py:113-115) runs on the GPU (kernel K0)."""
    def extractor(frame: np.ndarray, _unused=None):
        boxes = list(box_fn(frame))[:5]
        crops = []
        for top, right, bottom, left in boxes:
            c = frame[top:bottom, left:right]
            if c.size > 0:
                crops.append(torch.from_numpy(np.ascontiguousarray(c)).to(model._device))
        if not crops:
            return [], 0
        out = model.preprocess_crops(crops, swap_rb=True)
        return out.cpu().numpy(), len(crops)
    return extractor


def predict(filename: str, mtcnn=None) -> float:
    """cvit_prediction.py:153-242 with the model call replaced by the engine."""
    import cv2
    if model is None:
        raise RuntimeError("configure(model_=...) first")
    extractor = _face_extractor or face_face_rec
    face_tensor_face_rec = np.zeros((30, 224, 224, 3), dtype=np.uint8)
    cap = cv2.VideoCapture(filename)
    length = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    frame_count = int(length * 0.1)
    frame_jump = 5
    start_frame_number = 0
    loop = 0
    count_face_rec = 0
    while cap.isOpened() and loop < frame_count:
        loop += 1
        success, frame = cap.read()
        cap.set(cv2.CAP_PROP_POS_FRAMES, start_frame_number)
        if success:
            face_rec, count = extractor(frame, face_tensor_face_rec)
            if len(face_rec) and count > 0:
                kontrol = count_face_rec + count
                for f in face_rec:
                    if count_face_rec <= kontrol and (count_face_rec < 29):
                        face_tensor_face_rec[count_face_rec] = f
                        count_face_rec += 1
            start_frame_number += frame_jump
    cap.release()
    return predict_crops(face_tensor_face_rec[:count_face_rec], filename)


def predict_crops(store_rec: np.ndarray, filename: str = "") -> float:
    """The model half of ``predict`` (cvit_prediction.py:202-242) on uint8 crops [n,224,224,3]."""
    if len(store_rec) == 0:                                   # :218-219
        return torch.tensor(0.5).item()
    dfdc_tensor = torch.from_numpy(np.ascontiguousarray(store_rec)).to(model._device)
    n = dfdc_tensor.shape[0]
    # slot = index inside the <=32 chunk, chunks [0:32],[32:64],[64:90]; frames >= 90 do not enter the score
    # (:226-238) — both rules are applied by ff_cvit_predict itself
    scores = model.predict_videos(dfdc_tensor, [0, n])
    decCViT = scores[0]
    if verbose:
        print('CViT', filename, "Prediction:", decCViT.item())
    return decCViT.item()


def non_empty(dfdc_tensor, df_len, lower_bound, upper_bound, flag):
    """cvit_prediction.py:245-255."""
    thrtw = df_len
    if df_len >= upper_bound:
        thrtw = upper_bound
    if flag is True:
        return dfdc_tensor[lower_bound:thrtw]
    elif flag is False:
        return dfdc_tensor
    return []


def pred_sig(dfdc_tensor: torch.Tensor) -> torch.Tensor:
    """cvit_prediction.py:258-259 — element-wise sigmoid after squeeze()."""
    return torch.sigmoid(dfdc_tensor.squeeze())


def pred_tensor(dfdc_tensor, pre_tensor):
    """cvit_prediction.py:262-263."""
    return torch.cat((dfdc_tensor, pre_tensor), 0)


def pre_process_prediction(y_pred: torch.Tensor) -> torch.Tensor:
    """cvit_prediction.py:266-281 evaluated by the engine's reduction kernel (K8)."""
    if model is None:
        raise RuntimeError("configure(model_=...) first")
    if y_pred.dim() != 2 or len(y_pred) <= 2:
        return torch.tensor(0.5)
    # y_pred already holds pred_sig outputs: the kernel averages the probabilities it is handed
    off = torch.tensor([0, len(y_pred)], dtype=torch.int32)
    return model.video_scores(y_pred, off, mode=FF_REDUCE_REFERENCE_PROBS)[0].cpu()


def real_or_fake(predictions_or_score):
    """cvit_prediction.py:284-292 / README: < 0.5 REAL, >= 0.5 FAKE."""
    if isinstance(predictions_or_score, (float, int)):
        return "REAL" if predictions_or_score < 0.5 else "FAKE"
    return ["REAL" if p < 0.5 else "FAKE" for p in predictions_or_score]


def run(filenames: Sequence[str], save_csv_path: Optional[str] = None, num_workers: int = 1):
    """The ``__main__`` block (cvit_prediction.py:300-343): predictions + the filename,label CSV."""
    start_time = perf_counter()
    predictions = predict_on_video(filenames, num_workers=num_workers)
    times = perf_counter() - start_time
    if save_csv_path:
        import pandas as pd
        pd.DataFrame({"filename": list(filenames), "label": predictions}).to_csv(save_csv_path, index=False)
    return predictions, times
