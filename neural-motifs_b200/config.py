"""Constants of the hot path, same names and values as the reference's config.py:35-61.
(The argparse `ModelConfig` flag system, config.py:63-197, is caller-side and out of scope.)"""
MODES = ('sgdet', 'sgcls', 'predcls')

BOX_SCALE = 1024   # config.py:35
IM_SCALE = 592     # config.py:36

BG_THRESH_HI = 0.5
BG_THRESH_LO = 0.0

RPN_POSITIVE_OVERLAP = 0.7
RPN_NEGATIVE_OVERLAP = 0.3

RPN_FG_FRACTION = 0.5
FG_FRACTION = 0.25
RPN_BATCHSIZE = 256
ROIS_PER_IMG = 256
REL_FG_FRACTION = 0.25
RELS_PER_IMG = 256
RELS_PER_IMG_REFINE = 64

BATCHNORM_MOMENTUM = 0.01
ANCHOR_SIZE = 16

ANCHOR_RATIOS = (0.23232838, 0.63365731, 1.28478321, 3.15089189)
ANCHOR_SCALES = (2.22152954, 4.12315647, 7.21692515, 12.60263013, 22.7102731)
