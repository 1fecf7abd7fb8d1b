// TEST INFRASTRUCTURE ONLY: the reference's map_util.h includes "ros/ros.h" and uses nothing from it.
#pragma once
