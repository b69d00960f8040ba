"""Mirror of pointrcnn/lib/datasets: KittiDataset file access and the EVAL / TEST branch of KittiRCNNDataset."""
