"""Route 1 of INTEGRATION.md: an OpenChem config that names the B200 encoder class.

Same structure and hyper-parameters as the reference's example_configs/logp_gcnn_config.py (Graph2Label, GraphCNNEncoder 5x128,
encoder_dim 128, OpenChemMLP [128, 1], MSELoss, Adam lr 5e-4, StepLR 15/0.8, batch 256, F0 = 33 = 6+6+8+2+11 one-hots); the
ONLY model change is the import of GraphCNNEncoder (line marked below). The data layer is synthetic logP-shaped graphs
(openchem_b200.synthetic.SyntheticGraphDataset, GraphDataset's sample layout) because RDKit is not part of this image; with
RDKit installed, swap the two dataset lines for the reference's GraphDataset calls.

Run with the UNMODIFIED reference drivers (reference checkout on PYTHONPATH after this repo):
    python <OpenChem>/run.py --config_file=configs/logp_gcnn_b200.py --mode=train_eval
    python <OpenChem>/launch.py --nproc_per_node=8 <OpenChem>/run.py --config_file=configs/logp_gcnn_b200.py --mode=train_eval
Environment knobs: OCB200_DATA=logp_csv (the reference's logP_labels.csv under $OPENCHEM_ROOT instead of synthetic graphs),
OCB200_TRAIN / OCB200_VAL (synthetic data set sizes), OCB200_EPOCHS, OCB200_LOGDIR, OCB200_PRECISION,
OCB200_NMAX, OCB200_BATCH.
"""
import os

from openchem.models.Graph2Label import Graph2Label
from openchem_b200.modules.encoders.gcn_encoder import GraphCNNEncoder      # <- reference: openchem.modules.encoders.gcn_encoder
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
        'precision': os.environ.get("OCB200_PRECISION", "fp32"),
    },
    'mlp': OpenChemMLP,
    'mlp_params': {
        'input_size': 128,
        'n_layers': 2,
        'hidden_size': [128, 1],
        'activation': [F.relu, identity]
    }
}
