"""Prompt / image helpers the Stable-Diffusion driver imports (interface of the reference's qdiff_control/coco_prompt.py).
Data utilities only -- not on the quantized hot path."""
import json
import os
from random import shuffle


def get_prompts(json_file='/dataset/coco2014/annotations/captions_val2014.json'):
    with open(json_file, 'r') as f:
        data = json.load(f)
    prompts = [ann['caption'] for ann in data['annotations']]
    shuffle(prompts)
    return prompts


def center_resize_image(path_image, out_path, size):
    from PIL import Image  # only needed by the evaluation tooling
    os.makedirs(out_path, exist_ok=True)
    for filename in os.listdir(path_image):
        if not filename.lower().endswith(('.jpg', '.jpeg', '.png')):
            continue
        img = Image.open(os.path.join(path_image, filename))
        if filename.endswith('.JPEG') and img.mode == 'RGBA':
            continue
        width, height = img.size
        square = min(width, height)
        x1, y1 = (width - square) // 2, (height - square) // 2
        img.crop((x1, y1, x1 + square, y1 + square)).resize(size, resample=Image.BICUBIC).save(os.path.join(out_path, filename))
