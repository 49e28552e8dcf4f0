import sys
src = open('oracle/mg_experiments.py').read().replace('if __name__ == "__main__":', 'if False:').replace('if __name__ == "__main__" and len(sys.argv) > 4:', 'if True:')
exec(compile(src, 'mgx', 'exec'), {'__name__':'mgx','__file__':'oracle/mg_experiments.py'})
