from .pipeline import PipelineParallel, PipeSequential
