"""Run by tests/test_host.py through tools/emu/run_emulated.py: a CG that is cut short after two
iterations must hand over to the LU fallback (directFallback, the default) and still return the
solution; with directFallback=False it raises instead of returning the unconverged iterate."""
import warnings

import numpy

from icqsol_b200.mesh.generators import uvSphere
from icqsol_b200.mesh.polydata import PolyData
from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver

for storage in ('lower', 'full'):
    xyz, tri = uvSphere(1.0, 16, 9)
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage,
                           preconditioner='diagonal')
    solver.setSolverMaxNumberOfIterations(2)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        rsp = solver.computeResponseField()
    assert any('solving by LU' in str(w.message) for w in caught), [str(w.message) for w in caught]
    assert solver.lastSolveInfo['method'] == 'lu' and solver.lastSolveInfo['converged']
    g = solver.getGreenMatrix()
    assert numpy.abs(v + g.dot(rsp)).max() < 1e-11 * numpy.abs(v).max()
    assert solver.lastSolveInfo['cg']['iterations'] == 2 and not solver.lastSolveInfo['cg']['converged']
    # without the fallback the same cut-short solve raises: no unconverged iterate comes back
    plain = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage,
                          preconditioner='diagonal', directFallback=False)
    plain.setSolverMaxNumberOfIterations(2)
    plain.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    try:
        plain.computeResponseField()
        raise AssertionError('an unconverged solve must raise')
    except RuntimeError as e:
        assert 'CG stopped' in str(e)
    assert not plain.lastSolveInfo['converged']
    print('direct fallback ok:', storage)
