"""Dataset tables that DEFINE the text anchors and the evaluation loop of the reference (dataset/constants.py:1-147):
which classes a dataset has, the descriptive "real name" each class is prompted with, the domain that selects the
blur of calculate_similarity_map, and the prompt ensemble. This is data, reproduced value for value (any change would
change the anchors); it is stored here as one row per class and expanded into the reference's dictionaries below.
"""
from __future__ import annotations

import os
from typing import Dict, List

# (dataset, domain, data directory under AACLIP_DATA_ROOT, [(class, real name), ...]) in the reference's class order
_TABLE = [
    ("Brain", "Medical", "data/MedAD/Brain_AD", [("Brain", "scan")]),
    ("Liver", "Medical", "data/MedAD/Liver_AD", [("Liver", "scan")]),
    ("Retina", "Medical", "data/MedAD/Retina_RESC_AD", [("Retina", "scan")]),
    ("Colon_clinicDB", "Medical", "data/Colon/CVC-ClinicDB", [("Colon_clinicDB", "colon endoscopy image")]),
    ("Colon_colonDB", "Medical", "data/Colon/CVC-ColonDB", [("Colon_colonDB", "colon endoscopy image")]),
    ("Colon_Kvasir", "Medical", "data/Colon/Kvasir", [("Kvasir", "colon endoscopy image")]),
    ("Colon_cvc300", "Medical", "data/Colon/CVC-300", [("CVC-300", "colon endoscopy image")]),
    ("MVTec", "Industrial", "data/MVTec-AD", [
        ("bottle", "dark bottle"),
        ("cable", "top view of three cables"),
        ("capsule", "black and orange capsule"),
        ("carpet", "gray carpet"),
        ("grid", "metal or plastic mesh"),
        ("hazelnut", "single brown hazelnut"),
        ("leather", "brown leather"),
        ("metal_nut", "metal nut which has four notched edges"),
        ("pill", "oval white pill with small red speckles and the letters 'FF' engraved"),
        ("screw", "screw"),
        ("tile", "speckled tile surface"),
        ("transistor", "a three-legged transistor placed vertically"),
        ("toothbrush", "toothbrush head"),
        ("wood", "wood surface"),
        ("zipper", "a black zipper"),
    ]),
    ("VisA", "Industrial", "data/VisA_20220922", [
        ("candle", "candle"),
        ("pcb3", "infrared sensor pcb module"),
        ("capsules", "capsules"),
        ("pipe_fryum", "pipe-shaped fryum"),
        ("pcb4", "battery charging pcb module"),
        ("macaroni2", "scattered yellow macaroni"),
        ("pcb2", "integrated circuits board"),
        ("chewinggum", "chewing gum"),
        ("macaroni1", "orange macaroni"),
        ("cashew", "cashew nut"),
        ("fryum", "wheel-shaped fryum snack"),
        ("pcb1", "dual ultrasonic distance sensor pcb module"),
    ]),
    ("MPDD", "Industrial", "data/MPDD", [
        ("connector", "metal clamps with black adjustment knobs"),
        ("tubes", "scattered metal objects"),
        ("metal_plate", "blue rectangular metal plate with a notch on one side"),
        ("bracket_white", "white, elongated triangular metal bracket with a smooth, matte finish"),
        ("bracket_brown",
         "brown L-shaped metal bracket with smooth, glossy finish and multiple mounting holes along its arms"),
        ("bracket_black", "black ornamental metal bracket with spiral design attached to a rectangular frame"),
    ]),
    ("BTAD", "Industrial", "data/BTech_Dataset_transformed", [
        ("01", "Bright concentric rings in neon yellow and blue tones against a dark blue background, resembling a "
               "stylized wave or energy field radiating outward."),
        ("02", "vertical fabric lines in warm, dusty pink and beige tones"),
        ("03", "oval concentric circular rings in gradient shades of blue and white"),
    ]),
]

BASE_PATH = os.environ.get("AACLIP_DATA_ROOT", ".")
DATA_PATH: Dict[str, str] = {name: os.path.join(BASE_PATH, sub) for name, _, sub, _ in _TABLE}
CLASS_NAMES: Dict[str, List[str]] = {name: [c for c, _ in rows] for name, _, _, rows in _TABLE}
DOMAINS: Dict[str, str] = {name: dom for name, dom, _, _ in _TABLE}
REAL_NAMES: Dict[str, Dict[str, str]] = {name: dict(rows) for name, _, _, rows in _TABLE}

# the prompt ensemble (dataset/constants.py:137-147): state phrases x templates
PROMPTS = {
    "prompt_normal": ["{}", "a {}", "the {}"],
    "prompt_abnormal": ["a damaged {}", "a broken {}", "a {} with flaw", "a {} with defect", "a {} with damage"],
    "prompt_templates": ["{}.", "a photo of {}."],
}


def class_prompts(real_name: str):
    """The 6 normal + 10 abnormal sentences of one class, in the reference's order (forward_utils.py:148-153)."""
    groups = []
    for states in (PROMPTS["prompt_normal"], PROMPTS["prompt_abnormal"]):
        groups.append([t.format(s.format(real_name)) for s in states for t in PROMPTS["prompt_templates"]])
    return groups
