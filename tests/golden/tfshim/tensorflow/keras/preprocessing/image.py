class ImageDataGenerator:       # imported by sagan/dataset.py:7, used only by its image-folder loader
    def __init__(self, *a, **k):
        raise NotImplementedError
