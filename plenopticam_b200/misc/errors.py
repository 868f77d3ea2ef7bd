"""Stand-in for misc/errors.py:27-60: an exception that also appends to <exp_path>/err_log.txt."""
import datetime
import os


class PlenopticamError(Exception):

    def __init__(self, *args, **kwargs):
        super(PlenopticamError, self).__init__(*args)
        self.cfg = kwargs.get('cfg')
        self.sta = kwargs.get('sta')
        try:
            folder = self.cfg.exp_path if hasattr(self.cfg, 'exp_path') else os.path.expanduser("~")
            if self.sta is not None:
                self.sta.status_msg('Error! See log file in %s.' % os.path.join(folder, 'err_log.txt'))
            if os.path.isdir(folder):
                with open(os.path.join(folder, 'err_log.txt'), 'a') as f:
                    f.write('%s\n%s\n\n' % (datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S'), str(args)))
        except OSError:
            pass
