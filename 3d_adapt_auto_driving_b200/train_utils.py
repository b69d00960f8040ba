"""The one function of pointrcnn/tools/train_utils/train_utils.py that inference needs:
load_checkpoint (:78-92).  Checkpoints are {'epoch', 'it', 'model_state', 'optimizer_state'} dicts
(train_utils.py:60-75); state-dict keys are the reference's (tests/test_state_dict_compat.py)."""
import logging
import os

import torch

cur_logger = logging.getLogger(__name__)


def checkpoint_state(model=None, optimizer=None, epoch=None, it=None):
    model_state = None
    if model is not None:
        model_state = model.module.state_dict() if isinstance(model, torch.nn.DataParallel) else model.state_dict()
    return {'epoch': epoch, 'it': it, 'model_state': model_state,
            'optimizer_state': optimizer.state_dict() if optimizer is not None else None}


def save_checkpoint(state, filename='checkpoint'):
    torch.save(state, '{}.pth'.format(filename))


def load_checkpoint(model=None, optimizer=None, filename='checkpoint', logger=cur_logger):
    if not os.path.isfile(filename):
        raise FileNotFoundError
    logger.info("==> Loading from checkpoint '{}'".format(filename))
    checkpoint = torch.load(filename, map_location='cpu')
    epoch = checkpoint['epoch'] if 'epoch' in checkpoint.keys() else -1
    it = checkpoint.get('it', 0.0)
    if model is not None and checkpoint['model_state'] is not None:
        model.load_state_dict(checkpoint['model_state'])
    if optimizer is not None and checkpoint['optimizer_state'] is not None:
        optimizer.load_state_dict(checkpoint['optimizer_state'])
    logger.info("==> Done")
    return it, epoch
