#!/usr/bin/env python
"""
Solve Laplace equation given Dirichlet boundary conditions

Same command line as the reference's examples/solveLaplace.py (--input, --dirichlet,
--refine, --input_name, --output_name, --ascii, --output, --verbose); the matrix assembly
and the solve run on the GPU (icqsol_b200), the mesh is read and written with the
VTK-free readers (legacy POLYDATA, PLY) and the legacy POLYDATA writer.

    python examples/solveLaplace.py --input sphere.vtk --dirichlet 'sin(pi*x)*cos(pi*y)*z' \
        --output out.vtk --verbose
"""
from __future__ import print_function
import argparse
import os
import re
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver  # noqa: E402
from icqsol_b200.mesh.vtk_legacy_io import readVtkPolyData, writeVtkPolyData  # noqa: E402
from icqsol_b200.mesh.ply_io import readPly  # noqa: E402


def main(argv=None):
    # time stamp
    tid = re.sub(r'\.', '', str(time.time()))

    description = 'Solve Laplace equation given Dirichlet boundary conditions'
    parser = argparse.ArgumentParser(description=description)
    parser.add_argument('--input', dest='input', default='',
                        help='Input file (PLY or legacy VTK POLYDATA).')
    parser.add_argument('--dirichlet', dest='dirichlet', default='sin(pi*x)*cos(pi*y)*z',
                        help='Dirichlet boundary conditions, expression of x, y, and z.')
    parser.add_argument('--refine', dest='refine', default=0.0, type=float,
                        help='Maximum edge length (use 0 if no refinement).')
    parser.add_argument('--input_name', dest='input_name', default='voltage',
                        help='Set the name of the input field.')
    parser.add_argument('--output_name', dest='output_name', default='normal_electric_field',
                        help='Set the name of the output field.')
    parser.add_argument('--ascii', dest='ascii', action='store_true',
                        help='Save data in ASCII format (default is binary).')
    parser.add_argument('--output', dest='output', default='solveLaplace-{0}.vtk'.format(tid),
                        help='VTK Output file.')
    parser.add_argument('--verbose', dest='verbose', action='store_true',
                        help='Print info messages.')
    args = parser.parse_args(argv)

    if not args.input:
        print('ERROR: must specify input file: --input <file>')
        return 3

    # make sure the field names contain no spaces
    args.input_name = re.sub(r'\s', '_', args.input_name)
    args.output_name = re.sub(r'\s', '_', args.output_name)

    # the format of the input -- either vtk or ply, by suffix as util.getFileFormat does
    # (reference examples/solveLaplace.py:58-69)
    if args.input.lower().endswith('.ply'):
        pdata = readPly(args.input)
    else:
        pdata = readVtkPolyData(args.input)

    maxEdgeLength = float('inf')
    if args.refine > 0:
        maxEdgeLength = args.refine

    solver = LaplaceSolver(pdata, maxEdgeLength)

    # Set the input field.
    solver.setSourceFieldName(args.input_name)
    solver.setSourceFromExpression(args.dirichlet)
    solver.setResponseFieldName(args.output_name)

    # In place operation, pdata will be modified.
    normalEJump = solver.computeResponseField()

    if args.verbose:
        minJump = min(normalEJump)
        maxJump = max(normalEJump)
        avgJump = normalEJump.sum() / len(normalEJump)
        print('normal electric field jump min/avg/max: {0}/{1}/{2}'.format(minJump, avgJump,
                                                                           maxJump))
        # the potential given back by the response: G (-sigma) - v, through the device product
        potential = solver.getSourceArray(solver.getSourceArrayIndex())
        error = solver.applyGreenMatrix(-normalEJump) - potential
        print('total error: ', error.sum())
        info = solver.lastSolveInfo
        print('CG iterations: {0}, max|v + G sigma|: {1:.3e}, solve {2:.2f} ms, assembly {3:.2f} ms'.format(
            info['iterations'], info['residualInf'], info['milliseconds'],
            solver.assemblyMilliseconds['offDiagonal'] + solver.assemblyMilliseconds['diagonal']))

    if args.output:
        # Always produce VTK POLYDATA.
        writeVtkPolyData(solver.getVtkPolyData(), args.output, binary=not args.ascii)
    return 0


if __name__ == '__main__':
    sys.exit(main())
