// TEST INFRASTRUCTURE ONLY: faster/src/solverGurobi.cpp includes <ros/package.h> and uses nothing from it.
#pragma once
