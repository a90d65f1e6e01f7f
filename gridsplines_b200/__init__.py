"""gridsplines_b200 -- B200-native (sm_100a, FP64 CUDA) implicit heat/diffusion time-stepper of
gridsplines behind the reference's grid / Spline API.  See DESIGN.md and include/gs_b200.h."""
from ._lib import (CONST, CUBIC, DENSE, FLAG_NONFINITE, FLAG_NOT_DOMINANT, FLAG_SPLINE_UNDEFINED, FLOW, GRID,
                   INTERLEAVED, LEFT, LIB_PATH, LINE, LINE_MAJOR, LOW, MAP_DENSE, MAP_PER_COL, MAP_PER_ROW,
                   MAP_SINGLE, ORTHO, PREDICT, RADIAL, RESULT, RIGHT, SEL_APPLY, SEL_CAP, SEL_TEMPCOND, SEL_THERM,
                   SOLVER_AUTO, SOLVER_PARTITIONED, SOLVER_THOMAS, SPARSE, TEMP, UPP, GsError)
from .api import (AlwaysDefinedSpline, BatchedOneDGrid, RaggedLines, Coefficients, ConstantCoef, Context, DenseCoef, Flow, Many,
                  One, OneDGrid, Ortho, ShardedTwoDGrid, PatchedXCoef, PatchedYCoef, PatchXCoef, PatchYCoef, Radial, Spline, Temp, TwoDGrid,
                  VarXCoef, VarXCoefByColumn, VarYCoef, XDim, YDim,
                  default_context, discHeatFlow, discTWithFlow, getRadialMidPoint, radialHeatFlow)
