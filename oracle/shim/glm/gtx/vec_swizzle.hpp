// Empty on purpose: included by voxel.cuh:4, nothing from it is used by live code.
#pragma once
