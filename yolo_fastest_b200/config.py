"""Configuration dictionary with the reference's keys (src/model_training/_config.py:1-51) for the
parts the detection hot path reads (io_params).  ``config_for`` returns the per-resolution variant
the reference's comment describes at _config.py:5-9: 256x320 uses the first two anchor groups,
512x640 the last two."""
import copy

config_params = {
    "io_params": {
        "anchors": [
            [[10, 13], [16, 30], [33, 23]],
            [[150, 75], [100, 100], [75, 150]],
            [[300, 150], [200, 200], [150, 300]],
        ],
        "input_channel": 1,
        "input_shape": [256, 320, 1],
        "origin_img_shape": [512, 640, 3],
        "input_tensor_shape": (1, 1, 256, 320),
        "num_cls": 3,
        "num_anchors": 3,
        "anchor_mask": [[0, 1, 2], [3, 4, 5]],
        "strides": [16, 32],
        "conf_thre": 0.5,
        "nms_thre": 0.2,
        "class_names": ['carrier', 'defender', 'destroyer'],
    },
}

# anchors of the 80-class, 416x416 configuration named in yolo_fastest.py:403-406
COCO_ANCHORS = [[[12, 18], [37, 49], [52, 132]], [[115, 73], [119, 199], [242, 238]]]


def config_for(resolution):
    """resolution: "256x320" or "512x640" -> deep copy of config_params with input_shape and the two
    anchor groups that resolution was trained with."""
    cfg = copy.deepcopy(config_params)
    io = cfg["io_params"]
    if resolution == "256x320":
        io["input_shape"] = [256, 320, 1]
        io["anchors"] = io["anchors"][0:2]
    elif resolution == "512x640":
        io["input_shape"] = [512, 640, 1]
        io["anchors"] = io["anchors"][1:3]
    else:
        raise ValueError("unknown resolution %r" % (resolution,))
    io["input_tensor_shape"] = (1, 1, io["input_shape"][0], io["input_shape"][1])
    return cfg
