"""Stand-in for misc/status.py:27-170: progress/status observers and the cooperative `interrupt` Event.

The drop-in stages only use `status_msg`, `progress`, `interrupt` and `error`; the reference's own
`PlenopticamStatus` is accepted instead (duck typing).
"""
import threading


class PlenopticamStatus(object):

    def __init__(self):
        self._interrupt = threading.Event()
        self._error = False
        self._observers = {'prog': [], 'stat': [], 'interrupt': []}
        self._prog_var = 0
        self._stat_var = ''
        self.prog_opt = True

    # observer plumbing ------------------------------------------------------------------------------
    def bind_to_prog(self, callback):
        self._observers['prog'].append(callback)

    def bind_to_stat(self, callback):
        self._observers['stat'].append(callback)

    def bind_to_interrupt(self, callback):
        self._observers['interrupt'].append(callback)

    @property
    def prog_var(self):
        return self._prog_var

    @prog_var.setter
    def prog_var(self, val):
        self._prog_var = val
        for cb in self._observers['prog']:
            cb(val)

    @property
    def stat_var(self):
        return self._stat_var

    @stat_var.setter
    def stat_var(self, msg):
        self._stat_var = msg
        for cb in self._observers['stat']:
            cb(msg)

    # messages ---------------------------------------------------------------------------------------
    def progress(self, x, opt=False):
        if self.interrupt:
            self.prog_var = ''
        elif x is None:
            self.prog_var = 'Processing'
        elif isinstance(x, (int, float)):
            self.prog_var = 'Finished' if x == 100 else int(round(x, 0))
            if opt and self.prog_opt:
                print('\r Progress: {:2.0f}%'.format(float(x)), end='' if x != 100 else ' Finished\n')
        else:
            self.prog_var = str(x)
        return True

    def status_msg(self, msg='', opt=False):
        if self.interrupt:
            if not self._error:
                self.stat_var = 'Canceling'
            self.progress('', opt=False)
            return True
        self.stat_var = msg
        self.progress(None, opt=False)
        if opt:
            print('\n', msg)

    # interrupt / error ------------------------------------------------------------------------------
    @property
    def interrupt(self):
        return self._interrupt.is_set()

    @interrupt.setter
    def interrupt(self, val):
        if val:
            self._interrupt.set()
            for cb in self._observers['interrupt']:
                cb()
            self.status_msg()
        else:
            self._interrupt.clear()

    @property
    def error(self):
        return self._error

    @error.setter
    def error(self, val):
        self._error = bool(val)
        if val:
            self.interrupt = True
