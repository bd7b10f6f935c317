class VariableAggregation:
    MEAN = "mean"
