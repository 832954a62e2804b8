"""Answer emission of the eval scripts (host side; 002_train_vqa_arch1/004_eval_model.lua:236-273 and
004_eval_model_lf.lua:110-134): OpenEnded / MultipleChoice result JSON in the VQA evaluation format, and the late-fusion
weighted score sum.  The HDF5 inputs of the reference (data_prepro.h5 / data_img.h5) are not read here: no HDF5 library
exists in the image; callers pass arrays.  Plumbing only -- argmax and candidate selection come from the library
(``nvqa_argmax_get`` / ``nvqa_mc_select``)."""
import json

import numpy as np

from . import api


def late_fusion_scores(scores_a, scores_b, weight_a=0.5, weight_b=0.5):
    """004_eval_model_lf.lua:110-134: scores = weight_vgg * vgg_preds + weight_inception * inception_preds."""
    a, b = np.asarray(scores_a, dtype=np.float32), np.asarray(scores_b, dtype=np.float32)
    return (np.float32(weight_a) * a + np.float32(weight_b) * b).astype(np.float64)     # scores:double()


def open_ended_response(question_ids, pred, ix_to_ans):
    """004_eval_model.lua:248-251: [{question_id=qids[i], answer=ix_to_ans[tostring(pred[i])]}]; pred is 1-based."""
    return [{"question_id": int(q), "answer": ix_to_ans[str(int(p))]} for q, p in zip(question_ids, pred)]


def multiple_choice_response(question_ids, scores, mc_ids, ix_to_ans):
    """004_eval_model.lua:257-271: per question the best-scoring non-zero candidate id of MC_ans_test."""
    best = api.mc_select(np.asarray(scores, dtype=np.float32), np.asarray(mc_ids, dtype=np.int32))
    return [{"question_id": int(q), "answer": ix_to_ans[str(int(p))]} for q, p in zip(question_ids, best)]


def write_results(path_open_ended, path_multiple_choice, question_ids, scores, ix_to_ans, mc_ids=None, pred=None):
    """saveJson of both result files (004_eval_model.lua:253-255,273).  ``pred`` defaults to the first-max argmax of
    ``scores`` (torch.max semantics); pass the library's ``argmax`` output to avoid recomputing it."""
    scores = np.asarray(scores)
    if pred is None:
        pred = scores.argmax(axis=1) + 1                    # np.argmax returns the FIRST maximum, like torch.max
    with open(path_open_ended, "w") as f:
        json.dump(open_ended_response(question_ids, pred, ix_to_ans), f)
    if mc_ids is not None and path_multiple_choice:
        with open(path_multiple_choice, "w") as f:
            json.dump(multiple_choice_response(question_ids, scores, mc_ids, ix_to_ans), f)
