from .common import *  # noqa: F401,F403
from .common import validate_configuration, merge_configurations, zip_files, unzip_files, files_exist  # noqa: F401
