class Sequence:  # base class of the reference dataloaders (dataloader.py:19)
    pass
