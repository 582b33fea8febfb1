"""graphnet_classifier_b200 - B200-native (sm_100a) hot path of GraphNet_Classifier.

Mirrors the reference's module layout for the path in scope:

    reference                               here
    models/GNN.py, models/MLP.py            graphnet_classifier_b200.models.{GNN, MLP}
    utils/image_to_graph/*.py               graphnet_classifier_b200.utils.image_to_graph.*
    utils/dataloader.py                     graphnet_classifier_b200.utils.dataloader
    utils/train_model.py                    graphnet_classifier_b200.utils.train_model
    main.py (train_GNN)                     graphnet_classifier_b200.main

All arithmetic runs in libgnc.so (include/gnc.h, csrc/*.cu); importing the package
does not require a GPU, calling an operator does.
"""
__version__ = "0.1.0"
