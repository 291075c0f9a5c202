// stand-in (TEST INFRASTRUCTURE ONLY): the plugin registration is not exercised
#ifndef PLUGINLIB_EXPORT_CLASS
#define PLUGINLIB_EXPORT_CLASS(class_type, base_class_type)
#endif
