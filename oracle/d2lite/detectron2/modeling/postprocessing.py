from detectron2.structures import Instances


def detector_postprocess(results, output_height, output_width, mask_threshold=0.5):
    new_size = (int(output_height), int(output_width))
    scale_x, scale_y = output_width / results.image_size[1], output_height / results.image_size[0]
    results = Instances(new_size, **results.get_fields())
    output_boxes = results.pred_boxes if results.has("pred_boxes") else results.proposal_boxes
    output_boxes.scale(scale_x, scale_y)
    output_boxes.clip(results.image_size)
    return results[output_boxes.nonempty()]
