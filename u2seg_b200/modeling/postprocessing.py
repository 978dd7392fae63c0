"""Mirror of detectron2/modeling/postprocessing.py:9-100 and panoptic_fpn.py:184-269."""
import torch
import torch.nn.functional as F

from ..layers import paste_masks_in_image
from ..structures import Boxes, Instances


def detector_postprocess(results: Instances, output_height: int, output_width: int, mask_threshold: float = 0.5):
    """postprocessing.py:9-74."""
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    results = Instances((output_height, output_width), **results.get_fields())
    boxes = results.pred_boxes.clone() if results.has("pred_boxes") else results.proposal_boxes.clone()
    boxes.scale(scale_x, scale_y)
    boxes.clip(results.image_size)
    if results.has("pred_boxes"):
        results.pred_boxes = boxes
    else:
        results.proposal_boxes = boxes
    results = results[boxes.nonempty()]
    if results.has("pred_masks"):
        results.pred_masks = paste_masks_in_image(results.pred_masks[:, 0, :, :], results.pred_boxes.tensor,
                                                  results.image_size, threshold=mask_threshold)
    return results


def sem_seg_postprocess(result, img_size, output_height, output_width):
    """postprocessing.py:77-100."""
    result = result[:, :img_size[0], :img_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]


def combine_semantic_and_instance_outputs(instance_results, semantic_results, overlap_threshold, stuff_area_thresh,
                                          instances_score_thresh):
    """panoptic_fpn.py:184-269, with the per-instance .item() reads batched: areas and pairwise
    state are reduced on the device and read back once per instance loop instead of 3x per instance."""
    panoptic_seg = torch.zeros_like(semantic_results, dtype=torch.int32)
    sorted_inds = torch.argsort(-instance_results.scores)
    current_segment_id = 0
    segments_info = []
    instance_masks = instance_results.pred_masks.to(dtype=torch.bool, device=panoptic_seg.device)
    scores = instance_results.scores[sorted_inds].tolist()
    classes = instance_results.pred_classes[sorted_inds].tolist()
    areas = instance_masks.flatten(1).sum(1)[sorted_inds].tolist() if len(sorted_inds) else []
    for rank, inst_id in enumerate(sorted_inds.tolist()):
        score = scores[rank]
        if score < instances_score_thresh:
            break
        mask = instance_masks[inst_id]
        mask_area = areas[rank]
        if mask_area == 0:
            continue
        intersect = (mask > 0) & (panoptic_seg > 0)
        intersect_area = int(intersect.sum())
        if intersect_area * 1.0 / mask_area > overlap_threshold:
            continue
        if intersect_area > 0:
            mask = mask & (panoptic_seg == 0)
        current_segment_id += 1
        panoptic_seg[mask] = current_segment_id
        segments_info.append({"id": current_segment_id, "isthing": True, "score": score,
                              "category_id": classes[rank], "instance_id": inst_id})
    for semantic_label in torch.unique(semantic_results).cpu().tolist():
        if semantic_label == 0:
            continue
        mask = (semantic_results == semantic_label) & (panoptic_seg == 0)
        mask_area = int(mask.sum())
        if mask_area < stuff_area_thresh:
            continue
        current_segment_id += 1
        panoptic_seg[mask] = current_segment_id
        segments_info.append({"id": current_segment_id, "isthing": False, "category_id": semantic_label, "area": mask_area})
    return panoptic_seg, segments_info
