"""Drop-in for the reference's `mvs/mvs_dl.py:MVS_Inference` (SURVEY.md §8b level b1, the caller of the path):
same constructor, same `run(data_folder, mvs_path)`, same default-checkpoint lookup and the same fire-and-forget
`os.system` launch (`mvs_dl.py:28-65`) -- of `python -m deep3d_aerial_b200.predict` instead of
`python mvs/mvs_cas/predict.py`, on one process per GPU when `gpus > 1`.

    from deep3d_aerial_b200.mvs_dl import MVS_Inference           # run.py:30 imports it from mvs.mvs_dl
    MVS_Inference(max_w, max_h, view_num, num_depth, min_interval, model_type, pretrain_weight, display).run(export, mvs)

Like upstream it is meant to be started from the root of the reference checkout (paths `mvs/mvs_cas/...` are
relative); `reference_root` points elsewhere if needed.
"""
from __future__ import annotations

import os
import sys


class MVS_Inference:
    def __init__(self, max_w, max_h, view_num=5, num_depth=384, min_interval=0.1, model_type='adamvs',
                 pretrain_weight=None, display_depth=False, gpus=1, reference_root='mvs/mvs_cas', feature_cache=0,
                 master_port=29511):
        self.max_w = max_w
        self.max_h = max_h
        self.view_num = view_num
        self.num_depth = num_depth
        self.min_interval = min_interval
        self.pretrain_weight = pretrain_weight
        self.display_depth = display_depth
        self.model_type = model_type.lower()
        self.gpus = int(gpus)
        self.reference_root = reference_root
        self.feature_cache = int(feature_cache)
        self.master_port = int(master_port)

    def default_weight(self):
        """The last `.ckpt` listed under `<reference_root>/checkpoints/<model>/whu_omvs` (mvs_dl.py:46, 52-56)."""
        path = os.path.join(self.reference_root, 'checkpoints', self.model_type, 'whu_omvs')
        found = None
        for fname in os.listdir(path):
            if os.path.splitext(fname)[-1] == '.ckpt':
                found = os.path.join(path, fname)
        return found

    def command(self, data_folder, mvs_path):
        if self.model_type not in ["casmvsnet", "ucsnet", "msrednet", "adamvs"]:
            raise Exception("{}? Not implemented yet!".format(self.model_type))
        weight = self.pretrain_weight if self.pretrain_weight is not None else self.default_weight()
        launcher = '{} -m deep3d_aerial_b200.predict'.format(sys.executable)
        if self.gpus > 1:
            launcher = ('{} -m torch.distributed.run --nnodes=1 --nproc-per-node {} --master-addr 127.0.0.1 '
                        '--master-port {} -m deep3d_aerial_b200.predict'.format(sys.executable, self.gpus, self.master_port))
        cmd = ('{} --reference_root={} --data_folder={} --output_folder={} --model={} --loadckpt={} --view_num={} '
               '--numdepth={} --max_w={} --max_h={} --min_interval={} --display={}'.format(
                   launcher, self.reference_root, data_folder, mvs_path, self.model_type, weight, self.view_num,
                   self.num_depth, self.max_w, self.max_h, self.min_interval, self.display_depth))
        if self.feature_cache > 0:
            cmd += ' --feature_cache={} --partition=contiguous'.format(self.feature_cache)
        return cmd

    def run(self, data_folder, mvs_path):
        if not os.path.exists(os.path.dirname(mvs_path)):
            os.mkdir(os.path.dirname(mvs_path))
        str_ = self.command(data_folder, mvs_path)
        print(str_)
        os.system(str_)
