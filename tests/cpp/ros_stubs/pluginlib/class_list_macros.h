#pragma once
#define PLUGINLIB_EXPORT_CLASS(class_type, base_class_type)
