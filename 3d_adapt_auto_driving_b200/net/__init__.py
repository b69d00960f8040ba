"""Model graph of the PointRCNN inference path (mirror of pointrcnn/lib/net/)."""
