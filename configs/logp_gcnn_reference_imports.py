"""Route 2 of INTEGRATION.md: a config whose model imports are VERBATIM the reference's (example_configs/logp_gcnn_config.py:1-3).

Which GraphCNNEncoder it gets depends only on sys.path: with this repo's `openchem/` namespace shadow ahead of the reference
(run through run_b200.py) `openchem.modules.encoders.gcn_encoder` resolves to the B200 class; with the reference alone it is the
reference's eager implementation (that is how tests/test_dropin_host.py and `bench.py --impl reference` run the true reference).
Data layer and knobs: see configs/logp_gcnn_b200.py.
"""
import os

from openchem.models.Graph2Label import Graph2Label
from openchem.modules.encoders.gcn_encoder import GraphCNNEncoder
from openchem.modules.mlp.openchem_mlp import OpenChemMLP
from openchem.utils.utils import identity
from openchem_b200.synthetic import SyntheticGraphDataset

import torch.nn as nn
import torch.nn.functional as F
from torch.optim import Adam
from torch.optim.lr_scheduler import StepLR
from sklearn.metrics import r2_score

if os.environ.get("OCB200_DATA", "synthetic") == "logp_csv":
    # the reference's own data set (benchmark_datasets/logp_dataset/logP_labels.csv), same 80/20 split as logp_gcnn_config.py:51-58,
    # featurised without RDKit (openchem_b200/smiles.py: exact topology, approximate valence / hybridisation)
    from openchem_b200.smiles import logp_csv_datasets
    _root = os.environ.get("OPENCHEM_ROOT", ".")
    train_dataset, test_dataset = logp_csv_datasets(os.path.join(_root, "benchmark_datasets", "logp_dataset", "logP_labels.csv"))
else:
    _n = int(os.environ.get("OCB200_NMAX", "85"))
    train_dataset = SyntheticGraphDataset(int(os.environ.get("OCB200_TRAIN", "11340")), _n, 33, seed=1)
    test_dataset = SyntheticGraphDataset(int(os.environ.get("OCB200_VAL", "2836")), _n, 33, seed=2)

model = Graph2Label

model_params = {
    'task': 'regression',
    'random_seed': 42,
    'use_clip_grad': False,
    'batch_size': int(os.environ.get("OCB200_BATCH", "256")),
    'num_epochs': int(os.environ.get("OCB200_EPOCHS", "101")),
    'logdir': os.environ.get("OCB200_LOGDIR", 'logs/logp_gcnn_b200_logs'),
    'print_every': int(os.environ.get("OCB200_PRINT_EVERY", "1")),
    'save_every': int(os.environ.get("OCB200_SAVE_EVERY", "1")),
    'train_data_layer': train_dataset,
    'val_data_layer': test_dataset,
    'eval_metrics': r2_score,
    'criterion': nn.MSELoss(),
    'optimizer': Adam,
    'optimizer_params': {
        'lr': 0.0005,
    },
    'lr_scheduler': StepLR,
    'lr_scheduler_params': {
        'step_size': 15,
        'gamma': 0.8
    },
    'encoder': GraphCNNEncoder,
    'encoder_params': {
        'input_size': train_dataset[0]["node_feature_matrix"].shape[1],
        'encoder_dim': 128,
        'n_layers': 5,
        'hidden_size': [128, 128, 128, 128, 128],
    },
    'mlp': OpenChemMLP,
    'mlp_params': {
        'input_size': 128,
        'n_layers': 2,
        'hidden_size': [128, 1],
        'activation': [F.relu, identity]
    }
}
